"""The vectorised index sampler consumes the numpy Generator exactly like the reference's
call-by-call sampler (utilities/useful_statistics.py:73-109): same samples, same final state."""
from collections import Counter

import numpy as np

from aslsd_b200 import sampling


def _reference_stratum(i, j, h_inf, h_sup, n_samples, kappa, varpi, n_events, rng):
    """Literal restatement of sample_double_ix_stratum (useful_statistics.py:73-109)."""
    if h_inf == h_sup:
        h = h_inf
        first = varpi[i][j][h]
        list_m = first + rng.choice(n_events[i] - first, n_samples, replace=False)
        return {'m': list_m, 'n': kappa[j][i][list_m] - (h - 1)}
    list_h = h_inf + rng.choice(h_sup + 1 - h_inf, size=n_samples)
    counted = Counter(list_h)
    out_m, out_n = [], []
    for h in counted.keys():
        n_h = counted[h]
        first = varpi[i][j][h]
        pop = n_events[i] - first
        m = first + rng.choice(pop, n_h, replace=not (n_h <= pop))
        out_m.append(m)
        out_n.append(kappa[j][i][m] - h + 1)
    return {'m': np.concatenate(out_m), 'n': np.concatenate(out_n)}


def test_vectorised_sampler_keeps_the_reference_stream():
    from oracle import time_ordering as OT
    rng0 = np.random.default_rng(4)
    T_f = 400.
    times = [np.sort(rng0.uniform(0, T_f, size=n)) for n in (3000, 1800)]
    kappa, varpi = OT.kappa_varpi_closed_form(times, T_f)
    n_events = np.array([len(t) for t in times])
    for (i, j) in ((0, 0), (0, 1), (1, 0)):
        h_max = int(kappa[j][i][n_events[i] - 1]) + 1
        for (h_inf, h_sup, n) in ((1, 1, 50), (2, 40, 200), (41, h_max, 500), (3, 6, 300),
                                  (h_max - 3, h_max, 30)):
            a, b = np.random.default_rng(99), np.random.default_rng(99)
            got = sampling.sample_double_ix_stratum(i, j, h_inf, h_sup, n, kappa, varpi, n_events, a)
            want = _reference_stratum(i, j, h_inf, h_sup, n, kappa, varpi, n_events, b)
            np.testing.assert_array_equal(got['m'], want['m'])
            np.testing.assert_array_equal(got['n'], want['n'])
            assert str(a.bit_generator.state) == str(b.bit_generator.state)
