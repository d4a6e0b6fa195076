"""Oracle for the event-index maps kappa / varpi / lag_sizes (test infrastructure).

Follows ``stats/events/time_ordering.py`` of the reference:
  * ``get_kappa_varpi``  -> time_ordering.py:9-84
  * ``get_lag_sizes``    -> time_ordering.py:87-119

Two restatements are given: a literal one (``bisect`` per event, the same
sequence of searches as the reference, used for small inputs) and a closed form
on ``numpy.searchsorted``, checked equal to the
literal one in ``tests/test_oracle_index_maps.py``.
"""
import bisect

import numpy as np


def kappa_varpi_literal(list_times, T_f):
    """Per-event bisect restatement (time_ordering.py:48-82).

    kappa[i][j][n] : index of the last type-i event strictly before t^j_n
                     (-1 if none); the extra terminal entry kappa[i][j][N_j]
                     is ``bisect_left(t^i, T_f, 0, N_i-1)`` *without* the -1
                     (time_ordering.py:64-65).
    varpi[i][j][h] : first m such that t^i_m has at least h type-j
                     predecessors, i.e. first position of ``h-1`` in
                     kappa[j][i] (time_ordering.py:68-82).
    """
    d = len(list_times)
    n_ev = [len(t) for t in list_times]
    kappa = [[None] * d for _ in range(d)]
    for i in range(d):
        for j in range(d):
            if i == j:
                kappa[i][j] = np.arange(n_ev[i] + 1, dtype=np.int64) - 1
                continue
            row = np.full(n_ev[j] + 1, -1, dtype=np.int64)
            pos = 0
            for n in range(n_ev[j]):
                lo = max(pos, 0)
                pos = bisect.bisect_left(list_times[i], list_times[j][n],
                                         lo, n_ev[i]) - 1
                row[n] = pos
            row[n_ev[j]] = bisect.bisect_left(list_times[i], T_f, 0,
                                              n_ev[i] - 1)
            kappa[i][j] = row
    varpi = [[None] * d for _ in range(d)]
    for i in range(d):
        for j in range(d):
            if i == j:
                varpi[i][j] = np.arange(n_ev[i], dtype=np.int64)
                continue
            src = kappa[j][i]
            n_lags = int(src[-2]) + 2
            row = np.zeros(n_lags, dtype=np.int64)
            pos = -1
            for h in range(1, n_lags):
                lo = max(0, pos)
                pos = bisect.bisect_left(src, h - 1, lo, len(src))
                row[h] = pos
            varpi[i][j] = row
    return kappa, varpi


def kappa_varpi_closed_form(list_times, T_f):
    """``searchsorted`` closed form of the same maps.

    kappa[i][j][:N_j] = searchsorted(t^i, t^j, 'left') - 1
    varpi[i][j][1:H+1] = searchsorted(t^i, t^j[:H], 'right'),
                         H = kappa[j][i][N_i - 1] + 1
    """
    d = len(list_times)
    n_ev = [len(t) for t in list_times]
    kappa = [[None] * d for _ in range(d)]
    for i in range(d):
        for j in range(d):
            if i == j:
                kappa[i][j] = np.arange(n_ev[i] + 1, dtype=np.int64) - 1
                continue
            row = np.empty(n_ev[j] + 1, dtype=np.int64)
            row[:n_ev[j]] = np.searchsorted(list_times[i], list_times[j],
                                            side='left') - 1
            row[n_ev[j]] = np.searchsorted(list_times[i][:max(n_ev[i] - 1, 0)],
                                           T_f, side='left')
            kappa[i][j] = row
    varpi = [[None] * d for _ in range(d)]
    for i in range(d):
        for j in range(d):
            if i == j:
                varpi[i][j] = np.arange(n_ev[i], dtype=np.int64)
                continue
            H = int(kappa[j][i][n_ev[i] - 1]) + 1
            row = np.zeros(H + 1, dtype=np.int64)
            row[1:] = np.searchsorted(list_times[i], list_times[j][:H],
                                      side='right')
            varpi[i][j] = row
    return kappa, varpi


def lag_sizes(list_times, kappa, varpi):
    """lag_sizes[i][j][h] = N_i - varpi[i][j][h] for h >= 1, [0] = 0
    (time_ordering.py:112-118).  Returned as int64 arrays rather than the
    reference's Python lists; values are identical."""
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


def check_kappa_property(list_times, kappa):
    """Property from stats/events/tests/time_ordering_test.py:23-39:
    t^j[kappa] < t^i_m <= t^j[kappa+1]."""
    d = len(list_times)
    for i in range(d):
        for j in range(d):
            if i == j:
                continue
            tj, ti = list_times[j], list_times[i]
            for m in range(len(ti)):
                q = int(kappa[j][i][m])
                if q >= 0 and not tj[q] < ti[m]:
                    return False
                if q + 1 < len(tj) and not ti[m] <= tj[q + 1]:
                    return False
    return True


def check_varpi_property(list_times, varpi):
    """Property from stats/events/tests/time_ordering_test.py:42-66:
    t^i[varpi_h] > t^j_{h-1} >= t^i[varpi_h - 1]."""
    d = len(list_times)
    for i in range(d):
        for j in range(d):
            if i == j:
                continue
            ti, tj = list_times[i], list_times[j]
            row = varpi[i][j]
            for h in range(1, len(row)):
                m = int(row[h])
                if not ti[m] > tj[h - 1]:
                    return False
                if m >= 1 and not tj[h - 1] >= ti[m - 1]:
                    return False
    return True
