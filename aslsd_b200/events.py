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

    def truncate(self, T_trunc, engine=None):
        """Every path cut at ``T_trunc`` (episodes.py:90-95; upstream calls ``truncate`` on the list
        of paths instead of on each path)."""
        return Episodes([path.truncate(T_trunc) for path in self.list_paths], T_f=T_trunc)

    def clone(self):
        return Episodes([path.clone() for path in self.list_paths], T_f=self.T_f)

    def get_events_rate(self, interval_bounds):
        """Sum over the episodes of the per-path rates (episodes.py:109-115)."""
        rate = [np.zeros(len(interval_bounds[i]) - 1) for i in range(self.d)]
        for path in self.list_paths:
            for i, r in enumerate(path.get_events_rate(interval_bounds)):
                rate[i] += r
        return rate

    _STATS = ('n_episodes', 'T_f', 'd', 'list_n_events', 'list_eta', 'average_n_events',
              'total_n_events', 'average_eta')

    def save(self, file, **kwargs):
        """``file_path_<q>`` per episode plus ``file_stats_attr`` (episodes.py:120-131)."""
        import pickle
        for q, path in enumerate(self.list_paths):
            path.save(file + '_path_' + str(q))
        with open(file + '_stats_attr', 'wb', **kwargs) as f:
            pickle.dump({a: getattr(self, a) for a in self._STATS}, f)

    def load(self, file, **kwargs):
        import pickle
        with open(file + '_stats_attr', 'rb') as f:
            for a, val in pickle.load(f).items():
                setattr(self, a, val)
        self.list_paths = []
        for q in range(self.n_episodes):
            path = ProcessPath(book_keeping=False)
            path.load(file + '_path_' + str(q))
            self.list_paths.append(path)


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
        if self.d == 0:
            # an empty path (``p = ProcessPath(); p.load(file)``, episodes.py:147): nothing to index,
            # no device context needed
            self.kappa, self.varpi, self.lag_sizes = [], [], []
            return
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

    # ---- list operations (process_path.py:110-232) ------------------------------------------------
    def clone(self):
        """Same events and book-keeping in a new container; arrays are shared, the device context
        too.  (Upstream passes an unknown ``d=`` keyword to the constructor at :122 and raises
        TypeError; this is what that call evidently means.)"""
        return ProcessPath(self.list_times, self.T_f, T_i=self.T_i, list_marks=self.list_marks,
                           list_times2end=self.list_times2end, kappa=self.kappa, varpi=self.varpi,
                           lag_sizes=self.lag_sizes, book_keeping=self.book_keeping, engine=self.engine)

    def truncate(self, T_trunc, engine=None):
        """Events strictly before ``T_trunc`` (:132-170).  The reference slices its index maps
        (which leaves kappa without its terminal entry) and rebuilds lag_sizes; here the maps of
        the shorter path are rebuilt by one batched searchsorted launch - the same values on every
        index the slices hold (tests/test_process_path_mirror.py).  The truncated path gets a device
        context of its own (``engine``, default: a new one on the device of this path's engine), so
        the full path stays resident."""
        if T_trunc == 0:
            return ProcessPath([np.array([]) for _ in range(self.d)], T_trunc, book_keeping=False)
        if T_trunc >= self.T_f:
            return self.clone()
        keep = [int(np.searchsorted(t, T_trunc, side='left')) for t in self.list_times]
        marks = None if self.list_marks is None else [m[:n] for m, n in zip(self.list_marks, keep)]
        if engine is None and self.engine is not None:
            engine = Engine(self.engine.device) if isinstance(self.engine, Engine) else type(self.engine)()
        return ProcessPath([t[:n] for t, n in zip(self.list_times, keep)], T_trunc, list_marks=marks,
                           book_keeping=self.book_keeping, engine=engine)

    def censor(self, del_ixs, book_keeping=False):
        """Path without the events ``del_ixs[i]`` of every type i (:186-199; upstream's call fails
        on the same ``d=`` keyword as clone)."""
        times, marks = [], None if self.list_marks is None else []
        for i in range(self.d):
            drop = del_ixs[i]
            gone = drop is not None and len(drop) > 0
            times.append(np.delete(self.list_times[i], drop) if gone else self.list_times[i].copy())
            if marks is not None:
                m = self.list_marks[i]
                marks.append(np.delete(m, drop, axis=0) if gone else m.copy())
        return ProcessPath(times, self.T_f, list_marks=marks, book_keeping=book_keeping)

    def get_counting_process(self):
        """counting_process[i](t) = number of type-i events at or before t, on [0, T_f] (:172-184)."""
        from scipy.interpolate import interp1d
        self.counting_process = []
        for t in self.list_times:
            knots = np.concatenate([[0.], t, [self.T_f]])
            counts = np.concatenate([np.arange(len(t) + 1), [len(t)]])
            self.counting_process.append(interp1d(knots, counts, kind='previous'))
        return self.counting_process

    def get_inter_arrival_times(self):
        return [t[1:] - t[:-1] for t in self.list_times]

    def get_next_event(self, t_ref):
        """First event of any type strictly after ``t_ref`` (scalar or array), as (time, dim, mark)
        records (:208-232, path_event.py)."""
        if np.ndim(t_ref) == 0:
            return self.get_next_event(np.array([t_ref]))[0]
        order = np.argsort(np.concatenate(self.list_times))
        times = np.concatenate(self.list_times)[order]
        dims = np.concatenate([np.full(len(t), float(i)) for i, t in enumerate(self.list_times)])[order]
        marks = None if self.list_marks is None else np.concatenate(self.list_marks)[order]
        out = []
        for ix in np.searchsorted(times, t_ref, side='right'):
            out.append(PathEvent(time=times[ix], dim=dims[ix], mark=None if marks is None else marks[ix]))
        return out

    # ---- statistics (:237-296) --------------------------------------------------------------------
    def get_events_rate(self, interval_bounds):
        """Events of type i per unit of time in the bins ``interval_bounds[i]`` (:243-250; upstream
        cannot run: it subtracts lists and reads a ``self.times`` that does not exist)."""
        rate = []
        for i in range(self.d):
            bounds = np.asarray(interval_bounds[i], dtype=float)
            rate.append(np.histogram(self.list_times[i], bounds)[0] / (bounds[1:] - bounds[:-1]))
        return rate

    def get_sample_grid(self, sample_period):
        return sample_period * np.arange(int(self.T_f / sample_period) + 1)

    def get_sample_hist(self, sample_period, grid=None):
        if grid is None:
            grid = self.get_sample_grid(sample_period)
        return np.array([np.histogram(t, bins=grid)[0] for t in self.list_times])

    def get_empirical_intensity(self, sample_period, grid=None, sample_hist=None):
        if sample_hist is None:
            sample_hist = self.get_sample_hist(sample_period, grid=grid)
        return [sample_hist[i] / sample_period for i in range(self.d)]

    def get_empirical_covariance(self, sample_period, grid=None, max_tau=None, sample_hist=None,
                                 verbose=False):
        """cov[i, j, lag] of the binned counts of type i, ``lag`` bins after those of type j,
        per unit of time (:271-296)."""
        d = self.d
        n_bins = int(self.T_f / sample_period)
        if sample_hist is None:
            sample_hist = self.get_sample_hist(sample_period, grid=grid)
        max_lag = n_bins - 1 if max_tau is None else int(max_tau / sample_period)
        cov = np.zeros((d, d, max_lag + 1))
        for lag in range(max_lag + 1):
            late, early = sample_hist[:, lag:], sample_hist[:, :n_bins - lag]
            for i in range(d):
                for j in range(d):
                    cov[i, j, lag] = np.mean(early[j] * late[i]) - np.mean(early[j]) * np.mean(late[i])
        return cov / sample_period

    # ---- serialization (:301-336): one pickled dict, the reference's keys --------------------------
    _SAVED = ('list_times', 'list_marks', 'T_f', 'd', 'n_events', 'eta', 'book_keeping')
    _SAVED_MAPS = ('list_times2end', 'kappa', 'varpi', 'lag_sizes')

    def save(self, file, **kwargs):
        import pickle
        attrs = {a: getattr(self, a) for a in self._SAVED}
        if self.book_keeping:
            attrs.update({a: getattr(self, a) for a in self._SAVED_MAPS})
        with open(file, 'wb', **kwargs) as f:
            pickle.dump(attrs, f)

    def load(self, file, **kwargs):
        import pickle
        with open(file, 'rb') as f:
            attrs = pickle.load(f)
        for a in self._SAVED:
            setattr(self, a, attrs[a])
        for a in self._SAVED_MAPS:
            setattr(self, a, attrs[a] if self.book_keeping else None)
        self.T_i = getattr(self, 'T_i', 0.)
        self.engine = None


class PathEvent:
    """One event of a path (stats/events/path_event.py)."""

    def __init__(self, time=0., dim=0, mark=None, state=None):
        self.time, self.dim, self.mark, self.state = time, dim, mark, state
