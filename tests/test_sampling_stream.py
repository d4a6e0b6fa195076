"""The vectorised index sampler consumes the numpy Generator exactly like the reference's
call-by-call sampler (utilities/useful_statistics.py:73-109): same samples, same final state."""
from collections import Counter

import numpy as np
import pytest

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


def test_native_choice_is_numpy_choice():
    """aslsd_rng_choice against Generator.choice: Floyd's algorithm, the tail shuffle for large
    samples, with and without replacement, odd generator phases (32-bit output buffer)."""
    import ctypes
    from aslsd_b200 import capi
    lib = capi.load_library()
    cases = [(10, 3), (1, 1), (5, 5), (1000, 1), (1000, 37), (9999, 500), (10001, 150), (10001, 201),
             (10001, 300), (50000, 999), (50000, 1001), (50000, 12000), (50000, 12600), (50000, 20000), (10001, 2400), (10001, 2600), (200000, 1), (123457, 12),
             (70000, 70000), (3, 0)]
    for seed in (1, 2):
        for pop, n in cases:
            for replace in (False, True):
                if replace and pop == 0:
                    continue
                a, b = np.random.default_rng(seed), np.random.default_rng(seed)
                if seed == 2:
                    a.integers(0, 7), b.integers(0, 7)      # leave half of a 64-bit draw buffered
                want = a.choice(pop, n, replace=replace)
                lib_g = sampling._native_state(b)
                assert lib_g is not None
                g = lib_g[1]
                out = np.empty(n, dtype=np.int64)
                rc = lib.aslsd_rng_choice(ctypes.byref(g), pop, n, int(replace), capi.ptr_i64(out))
                assert rc == 0
                sampling._store_state(b, g)
                np.testing.assert_array_equal(out, want, err_msg='pop %d n %d replace %s' % (pop, n, replace))
                assert str(a.bit_generator.state) == str(b.bit_generator.state)


def test_native_term_sampler_keeps_the_reference_stream():
    from oracle import time_ordering as OT
    rng0 = np.random.default_rng(11)
    T_f = 500.
    times = [np.sort(rng0.uniform(0, T_f, size=n)) for n in (20000, 12000)]
    kappa, varpi = OT.kappa_varpi_closed_form(times, T_f)
    n_events = np.array([len(t) for t in times])
    for (i, j) in ((0, 0), (0, 1), (1, 0), (1, 1)):
        h_max = int(kappa[j][i][n_events[i] - 1]) + 1
        strata = [(1, 1), (2, 2), (3, 3), (4, 10), (11, 60), (61, 400), (401, h_max)]
        n_samples = [300, 20, 12000 if i == 0 else 7000, 150, 400, 900, 2500]
        a, b = np.random.default_rng(5), np.random.default_rng(5)
        got = sampling.sample_double_ix_stratified(i, j, strata, n_samples, kappa, varpi, n_events, a)
        want = [_reference_stratum(i, j, lo, hi, n, kappa, varpi, n_events, b)
                for (lo, hi), n in zip(strata, n_samples)]
        for g_s, w_s in zip(got, want):
            np.testing.assert_array_equal(g_s['m'], w_s['m'])
            np.testing.assert_array_equal(g_s['n'], w_s['n'])
        assert str(a.bit_generator.state) == str(b.bit_generator.state)


def test_native_single_index_sampler_keeps_the_reference_stream():
    """sample_single_ix_stratified through the native choice: numpy's values and final state."""
    strata = [(0, 49), (50, 20049), (20050, 70049), (70050, 70050), (70051, 200000)]
    n_samples = [50, 300, 1500, 1, np.int64(40000)]
    for seed in (3, 4):
        a, b = np.random.default_rng(seed), np.random.default_rng(seed)
        if seed == 4:
            a.integers(0, 7), b.integers(0, 7)
        want = [lo + a.choice(hi - lo + 1, size=n, replace=False) for (lo, hi), n in zip(strata, n_samples)]
        got = sampling.sample_single_ix_stratified(strata, n_samples, b)
        for w, g in zip(want, got):
            np.testing.assert_array_equal(g, w)
        assert str(a.bit_generator.state) == str(b.bit_generator.state)
        # flat mode: the same draws in one array, with the counts per stratum
        c = np.random.default_rng(seed)
        if seed == 4:
            c.integers(0, 7)
        flat, counts = sampling.sample_single_ix_stratified(strata, n_samples, c, flat=True)
        np.testing.assert_array_equal(flat, np.concatenate(want))
        np.testing.assert_array_equal(counts, [len(w) for w in want])
        assert str(a.bit_generator.state) == str(c.bit_generator.state)
    # an empty stratum list, a zero-size sample, and a request larger than the population
    rng = np.random.default_rng(0)
    assert sampling.sample_single_ix_stratified([], [], rng) == []
    got = sampling.sample_single_ix_stratified([(5, 9), (10, 12)], [0, 3], rng)
    assert len(got[0]) == 0 and sorted(got[1]) == [10, 11, 12]
    with pytest.raises(ValueError):
        sampling.sample_single_ix_stratified([(5, 9)], [6], rng)
    with pytest.raises(ValueError):
        sampling.sample_single_ix_stratified([(9, 5)], [1], rng)


def test_flat_double_index_sampler_is_the_list_form():
    rng = np.random.default_rng(5)
    n_i, n_j = 4000, 3500
    t_i, t_j = np.sort(rng.uniform(0, 100., n_i)), np.sort(rng.uniform(0, 100., n_j))
    from oracle import time_ordering as OT
    kappa, varpi = OT.kappa_varpi_closed_form([t_i, t_j], 100.)
    strata = [(1, 1), (2, 2), (3, 3), (4, 40), (41, 900), (901, int(kappa[1][0][n_i - 1]) + 1)]
    n_samples = [30, 30, 30, 200, 200, 200]
    for native in (True, False):
        a, b = np.random.default_rng(8), np.random.default_rng(8)
        if not native:     # a generator the native sampler does not restate: the numpy path
            a, b = (np.random.Generator(np.random.Philox(8)) for _ in range(2))
        want = sampling.sample_double_ix_stratified(0, 1, strata, n_samples, kappa, varpi, [n_i, n_j], a)
        m, n, counts = sampling.sample_double_ix_stratified(0, 1, strata, n_samples, kappa, varpi, [n_i, n_j], b,
                                                            flat=True)
        np.testing.assert_array_equal(m, np.concatenate([w['m'] for w in want]))
        np.testing.assert_array_equal(n, np.concatenate([w['n'] for w in want]))
        np.testing.assert_array_equal(counts, n_samples)
        assert str(a.bit_generator.state) == str(b.bit_generator.state)
        w1 = sampling.sample_single_ix_stratified([(0, 99), (100, 4000)], [10, 50], a)
        f1, c1 = sampling.sample_single_ix_stratified([(0, 99), (100, 4000)], [10, 50], b, flat=True)
        np.testing.assert_array_equal(f1, np.concatenate(w1))
        np.testing.assert_array_equal(c1, [10, 50])
