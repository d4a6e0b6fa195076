"""Event-path container and index maps, computed on the B200.

Mirrors of the reference interfaces (same names and argument meaning):
  * ``get_kappa_varpi(list_times, T_f)``     stats/events/time_ordering.py:9-84
  * ``get_lag_sizes(list_times, kappa, varpi)``  time_ordering.py:87-119
  * ``ProcessPath(list_times, T_f, ..., kappa=, varpi=, lag_sizes=)``
                                             stats/events/process_path.py:63-106

``lag_sizes[i][j]`` is returned as an int64 array instead of the reference's
Python list of mixed int / np.int64: consumers only slice
and ``np.sum`` it (double_ix_stratification.py:63-70), and a list would cost
gigabytes of PyObjects at 10^6 events per type.
"""
import numpy as np

from aslsd_b200.engine import Engine

_default_engine = None


def default_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


def _path_key(data):
    return (id(data.list_times), float(data.T_f), tuple(len(t) for t in data.list_times))


def engine_for(data):
    """Device context for the events of ``data``: the path's own ``engine`` attribute if it has
    one, the process-wide default engine while that one is free or already holds this path, and
    otherwise a new engine that is kept on the path object.  Several live paths (the episodes of
    ``fit_episodes``, mhp.py:1806-2008) thus stay resident on the device side by side instead of
    being re-uploaded every time the fit loop moves from one path to the next."""
    own = getattr(data, 'engine', None)
    if own is not None:
        return own
    eng = default_engine()
    held = getattr(eng, '_events_key', None)
    if held is None or held == _path_key(data):
        return eng
    eng = Engine(eng.device if hasattr(eng, 'device') else 0)
    try:
        data.engine = eng
    except AttributeError:   # a path type with __slots__: fall back to sharing the default engine
        return default_engine()
    return eng


class Episodes:
    """Several paths of the same process observed independently (stats/events/episodes.py:13-80):
    the container ``fit_episodes`` iterates over."""

    def __init__(self, list_paths=None, T_f=1.):
        self.list_paths = [] if list_paths is None else list(list_paths)
        self.n_episodes = len(self.list_paths)
        self.T_f = T_f
        self.d = self.list_paths[0].d if self.n_episodes else 0
        self.list_n_events = np.zeros((self.d, self.n_episodes))
        self.list_eta = np.zeros((self.d, self.n_episodes))
        for ix, path in enumerate(self.list_paths):
            self.list_n_events[:, ix] = path.n_events
            self.list_eta[:, ix] = path.eta
        if self.n_episodes > 0:
            self.average_n_events = np.mean(self.list_n_events, axis=1)
            self.total_n_events = np.sum(self.list_n_events, axis=1, dtype=int)
            self.average_eta = np.mean(self.list_eta, axis=1)


def get_kappa_varpi(list_times, T_f, engine=None):
    """kappa[j][i][m]: index of the latest type-j event strictly before t^i_m;
    varpi[i][j][h]: first m such that t^i_m has at least h type-j predecessors."""
    engine = engine or default_engine()
    for t in list_times:
        if len(t) == 0:
            # the reference indexes kappa[j][i][-2] (time_ordering.py:72)
            raise IndexError('get_kappa_varpi: an event type has no events')
    engine.set_events(list_times, T_f)
    kappa, varpi, _ = engine.index_maps(want_lag_sizes=False)
    return kappa, varpi


def get_lag_sizes(list_times, kappa, varpi):
    """lag_sizes[i][j][h] = N_i - varpi[i][j][h] for h >= 1, [0] = 0.
    O(sum of lags) host arithmetic on arrays the caller already holds; the fused
    device version is ``Engine.index_maps`` (used by ProcessPath)."""
    d = len(list_times)
    out = [[None] * d for _ in range(d)]
    for i in range(d):
        n_i = len(list_times[i])
        for j in range(d):
            n_lags = int(kappa[j][i][n_i - 1]) + 2
            row = np.zeros(n_lags, dtype=np.int64)
            row[1:] = n_i - np.asarray(varpi[i][j][1:n_lags], dtype=np.int64)
            out[i][j] = row
    return out


class ProcessPath:
    """Finite path of a d-dimensional point process with its book-keeping
    (process_path.py:63-106); the index maps are built by one batched
    searchsorted launch unless they are passed in."""

    def __init__(self, list_times=None, T_f=1., T_i=0., list_marks=None,
                 n_events=None, eta=None, list_times2end=None, kappa=None,
                 varpi=None, lag_sizes=None, book_keeping=True, engine=None):
        if list_times is None:
            list_times = []
        self.T_f = T_f
        self.T_i = T_i
        self.d = len(list_times)
        self.list_times = [0. + np.array(list_times[i], dtype=np.float64)
                           for i in range(self.d)]
        self.list_marks = (None if list_marks is None
                           else [0 + np.array(m) for m in list_marks])
        self.n_events = np.array([len(L) for L in self.list_times])
        self.eta = 0. if T_f == 0. else self.n_events / T_f
        self.book_keeping = book_keeping
        self.engine = engine
        self.list_times2end = list_times2end
        self.kappa, self.varpi, self.lag_sizes = kappa, varpi, lag_sizes
        if book_keeping:
            if list_times2end is None:
                self.list_times2end = [T_f - L for L in self.list_times]
            if kappa is None or varpi is None or lag_sizes is None:
                self.get_book_keeping(only_missing=True)

    def get_book_keeping(self, only_missing=False):
        self.list_times2end = [self.T_f - L for L in self.list_times]
        for t in self.list_times:
            if len(t) == 0:
                raise IndexError('ProcessPath: an event type has no events')
        engine = self.engine or default_engine()
        engine.set_events(self.list_times, self.T_f)
        kappa, varpi, lag_sizes = engine.index_maps()
        if not only_missing or self.kappa is None or self.varpi is None:
            self.kappa, self.varpi = kappa, varpi
            self.lag_sizes = lag_sizes
        elif self.lag_sizes is None:
            self.lag_sizes = get_lag_sizes(self.list_times, self.kappa, self.varpi)
