"""Parity of the exact LSE sums computed by the CUDA pair kernels, through the
C ABI, against (1) the reference's golden vectors, (2) the oracle on seeded
inputs, (3) size-independent properties at larger sizes.

Tolerance: relative 1e-10 per term (BASELINE.json north_star), fp64."""
import numpy as np
import pytest

from cases import CASES, NAMES_2D, case_x
from helpers import RTOL, load_golden, golden_times, assert_sums_close, iter_sum_entries
from oracle import exact as OE

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
    import aslsd_b200
    return aslsd_b200.Engine(0)


@pytest.mark.parametrize('name', NAMES_2D)
def test_exact_sums_match_reference_golden(engine, name):
    model = CASES[name][0]
    g = load_golden('exact_' + name)
    times, T_f = golden_times(g)
    engine.set_events(times, T_f)
    engine.set_model(model)
    for k in range(model['d']):
        sum_f, sum_diff_f = engine.exact_sums(k, g['k%d/x' % k])
        worst = assert_sums_close(sum_f, sum_diff_f, g, 'k%d' % k, rtol=RTOL)
        assert worst <= RTOL


def test_benchmark_model_matches_reference_golden(engine):
    """The benchmark's own model (config C5: ten event types, SBF-Gaussian(10)) against the LIVE
    REFERENCE's sums of sub-problems 0, 4 and 9 on a short path (tests/golden/make_golden_c5model.py;
    every target of that path lies near T_f, so the per-item kernels evaluate it - the lag quadrature
    of the same model is checked against the oracle in test_c5_model_matches_oracle_every_k)."""
    from aslsd_b200 import simulate
    g = load_golden('exact_c5model')
    times, T_f = golden_times(g)
    engine.set_events(times, T_f)
    engine.set_model(simulate.sbf_gauss_model(10, 10))
    for k in (int(v) for v in g['ks']):
        sum_f, sum_diff_f = engine.exact_sums(k, g['k%d/x' % k])
        worst = assert_sums_close(sum_f, sum_diff_f, g, 'k%d' % k, rtol=RTOL)
        assert worst <= RTOL


@pytest.mark.parametrize('name', ['exp2d', 'sbf2d3', 'gauss2d'])
def test_function_only_flag(engine, name):
    model = CASES[name][0]
    g = load_golden('exact_' + name)
    times, T_f = golden_times(g)
    engine.set_events(times, T_f)
    engine.set_model(model)
    sum_f, _ = engine.exact_sums(0, g['k0/x'], compute_f_sum=True, compute_diff_f_sum=False)
    for name_t, val in sum_f.items():
        ref = g['k0/f/%s' % name_t]
        assert abs(val - ref) <= RTOL * abs(ref)


def _simulate_poisson_cluster(rng, d, T_f, rate):
    # plain superposition of homogeneous streams: enough for a parity input
    return [np.sort(rng.uniform(0, T_f, size=rng.poisson(rate * T_f))) for _ in range(d)]


@pytest.mark.parametrize('name', ['exp2d', 'gauss2d', 'sbf2d3', 'sbf1d10', 'exp1d2b', 'gauss1d2b'])
def test_exact_sums_match_windowed_oracle_medium(engine, name):
    """~10^4 events per type: the oracle's windowed restatement (agrees with the
    literal reference to ~1e-15, tests/test_oracle_golden.py)."""
    model = CASES[name][0]
    d = model['d']
    rng = np.random.default_rng(99)
    times = _simulate_poisson_cluster(rng, d, 4000., 2.0 if 'sbf' not in name else 1.0)
    T_f = 4000.
    engine.set_events(times, T_f)
    engine.set_model(model)
    xs = case_x(name)
    for k in range(d):
        sum_f, sum_diff_f, counts = engine.exact_sums(k, xs[k], count_pairs=True)
        o_f, o_df, o_counts = OE.exact_sums_windowed(model, k, xs[k], times, T_f)
        ref = {}
        for label, val in iter_sum_entries(o_f, o_df):
            ref['k/%s' % label] = val
        assert_sums_close(sum_f, sum_diff_f, ref, 'k', rtol=RTOL)


def test_shards_add_up_and_match_oracle_blocks(engine):
    """Time-block shards of the target events add up to the full sums."""
    name = 'sbf2d3'
    model = CASES[name][0]
    rng = np.random.default_rng(5)
    times = _simulate_poisson_cluster(rng, 2, 3000., 1.5)
    T_f = 3000.
    engine.set_events(times, T_f)
    engine.set_model(model)
    x = case_x(name)[1]
    full = engine.exact_sums_packed(1, x)
    cuts = [0.0, 700., 1800., T_f + 1]
    acc = np.zeros_like(full)
    for a, b in zip(cuts[:-1], cuts[1:]):
        ranges = [(int(np.searchsorted(t, a)), int(np.searchsorted(t, b))) for t in times]
        part = engine.exact_sums_packed(1, x, ranges=ranges)
        acc += part
        o_f, o_df, _ = OE.exact_sums_windowed(model, 1, x, times, T_f, ranges=ranges)
        from aslsd_b200.engine import unpack_terms
        s_f, s_df = unpack_terms(engine.desc, 1, part)
        ref = {'k/%s' % label: val for label, val in iter_sum_entries(o_f, o_df)}
        assert_sums_close(s_f, s_df, ref, 'k', rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(acc, full, rtol=1e-12, atol=1e-12)


def test_time_block_slices_with_halo_add_up(engine):
    """Each shard holds only its time block plus the halo of the largest lag
    (aslsd_exact_max_lag): the shard sums still add up to the full sums."""
    from aslsd_b200.distributed import block_slices, time_block_ranges
    name = 'sbf2d3'
    model = CASES[name][0]
    rng = np.random.default_rng(15)
    T_f = 4000.
    times = _simulate_poisson_cluster(rng, 2, T_f, 1.5)
    xs = case_x(name)
    engine.set_events(times, T_f)
    engine.set_model(model)
    full = engine.exact_sums_batch([0, 1], xs)
    max_lag = engine.exact_max_lag([0, 1], xs)
    assert 0. < max_lag < 40.
    acc = [np.zeros_like(f) for f in full]
    held = 0
    for block in time_block_ranges(times, 3):
        local, ranges = block_slices(times, block, max_lag)
        held += sum(len(t) for t in local)
        engine.set_events(local, T_f)
        for a, part in zip(acc, engine.exact_sums_batch([0, 1], xs, ranges=ranges)):
            a += part
    for a, f in zip(acc, full):
        np.testing.assert_allclose(a, f, rtol=1e-11, atol=1e-10)
    assert held < 1.1 * sum(len(t) for t in times)     # blocks + small halos, not 3 copies


def test_split_raw_and_assemble_matches_fused(engine):
    import torch
    name = 'gauss2d'
    model = CASES[name][0]
    g = load_golden('exact_' + name)
    times, T_f = golden_times(g)
    engine.set_events(times, T_f)
    engine.set_model(model)
    x = g['k1/x']
    fused = engine.exact_sums_packed(1, x)
    raw = torch.zeros(engine.exact_raw_len(1), dtype=torch.float64, device='cuda:0')
    engine.exact_raw_device(1, x, raw.data_ptr())
    torch.cuda.synchronize()
    out = engine.exact_assemble(1, x, raw.cpu().numpy())
    np.testing.assert_array_equal(out, fused)


@pytest.mark.parametrize('name', ['sbf2d3', 'gauss2d', 'exp2d'])
def test_batch_equals_per_k_and_shares_rows(engine, name):
    """aslsd_exact_sums_batch: same packed sums as one call per k; with (beta, delta)
    fixed the upsilon rows of the d sub-problems coincide and are evaluated once."""
    model = CASES[name][0]
    d = model['d']
    rng = np.random.default_rng(21)
    times = _simulate_poisson_cluster(rng, d, 3000., 1.5)
    engine.set_events(times, 3000.)
    engine.set_model(model)
    xs = case_x(name)
    singles = [engine.exact_sums_packed(k, xs[k]) for k in range(d)]
    outs, counts = engine.exact_sums_batch(list(range(d)), xs, count_pairs=True)
    work = engine.last_work_counts()
    for k in range(d):
        # the quadrature bins of a batch differ from those of a single k (other rows share the
        # streams): agreement to the quadrature's own accuracy, not bit for bit
        np.testing.assert_allclose(outs[k], singles[k], rtol=1e-11, atol=1e-11)
        _, c1 = engine.exact_sums_packed(k, xs[k], count_pairs=True)
        assert counts[k] == c1
    assert work['rows'] <= work['items']
    if name == 'sbf2d3':
        # per k: 2 x (3 psi + 9 upsilonzero + 3 phi + 9 upsilon_ii) + 2 x 9 cross = 66 items;
        # upsilon (36), upsilonzero (18) and psi (6) rows are shared by k = 0 and k = 1
        assert work['items'] == 132 and work['rows'] == 72
        assert work['moment_pairs'] > 0 and work['streams'] == 4


@pytest.mark.parametrize('name', NAMES_2D)
def test_lag_quadrature_agrees_with_item_kernels(engine, name):
    """Lag quadrature (interior targets of every double sum) against the per-item
    pair kernels on the same input, and both against the oracle."""
    model = CASES[name][0]
    d = model['d']
    rng = np.random.default_rng(17)
    times = _simulate_poisson_cluster(rng, d, 6000., 1.2)
    T_f = 6000.
    engine.set_events(times, T_f)
    engine.set_model(model)
    xs = case_x(name)
    from aslsd_b200.engine import unpack_terms
    try:
        for k in range(d):
            engine.use_lag_quadrature(True)
            a = engine.exact_sums_packed(k, xs[k])
            wa = engine.last_work_counts()
            engine.use_lag_quadrature(False)
            b = engine.exact_sums_packed(k, xs[k])
            wb = engine.last_work_counts()
            # term by term: |a - b| <= 1e-12 |b| (+ a floor for derivative sums that cancel to ~0)
            fa, dfa = unpack_terms(engine.desc, k, a)
            fb, dfb = unpack_terms(engine.desc, k, b)
            ref = {'k/%s' % label: val for label, val in iter_sum_entries(fb, dfb)}
            assert_sums_close(fa, dfa, ref, 'k', rtol=1e-12, atol=1e-11)
            assert wb['moment_pairs'] == 0 and wb['streams'] == 0
            assert wa['moment_pairs'] > 0 and wa['pair_evals'] < wb['pair_evals']
            o_f, o_df, _ = OE.exact_sums_windowed(model, k, xs[k], times, T_f)
            ref = {'k/%s' % label: val for label, val in iter_sum_entries(o_f, o_df)}
            assert_sums_close(fa, dfa, ref, 'k', rtol=RTOL)
    finally:
        engine.use_lag_quadrature(True)


@pytest.mark.parametrize('family', ['exp', 'gauss', 'mix'])
def test_lag_quadrature_random_parameters(engine, family):
    """Random (beta, delta) over two decades of scale, clustered event times with exact
    ties: quadrature against the per-item kernels."""
    rng = np.random.default_rng(2024)
    from aslsd_b200.engine import unpack_terms
    try:
        for trial in range(6):
            d = 2
            if family == 'exp':
                basis = [{'type': 'exp'}, {'type': 'exp'}]
                def draw():
                    return [0.3 * rng.uniform(), 10 ** rng.uniform(-0.5, 0.7),
                            0.3 * rng.uniform(), 10 ** rng.uniform(-0.5, 0.7)]
            elif family == 'gauss':
                basis = [{'type': 'gauss'}, {'type': 'gauss'}]
                def draw():
                    return [0.3 * rng.uniform(), 10 ** rng.uniform(-0.7, 0.3), rng.uniform(0., 4.),
                            0.3 * rng.uniform(), 10 ** rng.uniform(-0.7, 0.3), rng.uniform(0., 4.)]
            else:
                # exponential x Gaussian pairs in both orders; beta_exp * beta_gauss stays moderate
                # (beyond ~3 the reference's own cdf differences lose their digits)
                basis = [{'type': 'exp'}, {'type': 'gauss'}]
                def draw():
                    return [0.3 * rng.uniform(), 10 ** rng.uniform(-0.5, 0.3),
                            0.3 * rng.uniform(), 10 ** rng.uniform(-0.7, 0.0), rng.uniform(0., 4.)]
            model = {'d': d, 'kernels': [[basis for _ in range(d)] for _ in range(d)]}
            T_f = 3000.
            times = []
            for _ in range(d):
                parents = np.sort(rng.uniform(0, T_f, size=2500))
                kids = parents[rng.integers(0, len(parents), size=2500)] + rng.exponential(0.05, size=2500)
                t = np.sort(np.concatenate([parents, kids[kids < T_f], parents[:50]]))   # 50 exact ties
                times.append(t)
            engine.set_events(times, T_f)
            engine.set_model(model)
            for k in range(d):
                x = np.array([0.7] + draw() + draw())
                engine.use_lag_quadrature(True)
                a = engine.exact_sums_packed(k, x)
                wa = engine.last_work_counts()
                engine.use_lag_quadrature(False)
                b = engine.exact_sums_packed(k, x)
                fa, dfa = unpack_terms(engine.desc, k, a)
                fb, dfb = unpack_terms(engine.desc, k, b)
                ref = {'k/%s' % label: val for label, val in iter_sum_entries(fb, dfb)}
                assert_sums_close(fa, dfa, ref, 'k', rtol=1e-11, atol=1e-10)
                assert wa['moment_pairs'] > 0
    finally:
        engine.use_lag_quadrature(True)


def test_lag_quadrature_unbalanced_rates_and_many_bins(engine):
    """A slow target stream against a fast source stream (the source window of a tile does not
    fit the staged buffer: global-memory fallback of the moment kernel) and a wide spread of
    decay rates (hundreds of bins: 10-bit ranks, small sort chunks)."""
    rng = np.random.default_rng(77)
    from aslsd_b200.engine import unpack_terms
    T_f = 2500.
    times = [np.sort(rng.uniform(0, T_f, size=600)), np.sort(rng.uniform(0, T_f, size=24000))]
    model = {'d': 2, 'kernels': [[[{'type': 'exp'}] for _ in range(2)] for _ in range(2)]}
    engine.set_events(times, T_f)
    engine.set_model(model)
    try:
        for k, x in enumerate([np.array([0.5, 0.2, 0.3, 0.1, 3.0]), np.array([0.4, 0.1, 1.0, 0.2, 0.8])]):
            engine.use_lag_quadrature(True)
            a = engine.exact_sums_packed(k, x)
            wa = engine.last_work_counts()
            engine.use_lag_quadrature(False)
            b = engine.exact_sums_packed(k, x)
            fa, dfa = unpack_terms(engine.desc, k, a)
            fb, dfb = unpack_terms(engine.desc, k, b)
            ref = {'k/%s' % label: val for label, val in iter_sum_entries(fb, dfb)}
            assert_sums_close(fa, dfa, ref, 'k', rtol=1e-11, atol=1e-10)
            assert wa['moment_pairs'] > 0 and wa['bins'] > 128 * wa['streams'] / 2
            o_f, o_df, _ = OE.exact_sums_windowed(model, k, x, times, T_f)
            ref = {'k/%s' % label: val for label, val in iter_sum_entries(o_f, o_df)}
            assert_sums_close(fa, dfa, ref, 'k', rtol=RTOL, atol=1e-10)
    finally:
        engine.use_lag_quadrature(True)


@pytest.mark.parametrize('sigma,delta', [(1.0, 0.0), (0.4, 4.0)])
def test_warp_shape_dense_sources_bursts_and_ties(engine, sigma, delta):
    """Gaussian kernels (streams of at most 64 bins: the one-warp-per-work-unit moment kernel,
    one and two passes) on a sparse stream against a dense one with bursts of exact ties: source
    windows that do not fit the staged buffer (global-memory tiles), tiles of tens of thousands of
    pairs (several mark chunks), steps with more pairs of one bin than a stack holds (pair-by-pair
    path), partial drains - against the per-item kernels."""
    rng = np.random.default_rng(909)
    from aslsd_b200.engine import unpack_terms
    T_f = 300.
    sparse = np.sort(rng.uniform(0, T_f, size=700))
    dense = rng.uniform(0, T_f, size=30000)
    bursts = np.repeat(rng.uniform(5., T_f - 5., size=12), 40)          # 12 bursts of 40 simultaneous events
    dense = np.sort(np.concatenate([dense, bursts, sparse[:30]]))      # and ties across the two types
    times = [sparse, dense]
    basis = [{'type': 'gauss'}]
    model = {'d': 2, 'kernels': [[basis for _ in range(2)] for _ in range(2)]}
    engine.set_events(times, T_f)
    engine.set_model(model)
    engine.use_weight_cache(False)      # the two sub-problems share (beta, delta): every call bins its pairs
    try:
        for k in range(2):
            x = np.array([0.6, 0.2, sigma, delta, 0.1, 1.3 * sigma, 0.5 * delta])
            engine.use_lag_quadrature(True)
            a = engine.exact_sums_packed(k, x)
            wa = engine.last_work_counts()
            a2 = engine.exact_sums_packed(k, x)
            engine.use_lag_quadrature(False)
            b = engine.exact_sums_packed(k, x)
            assert wa['moment_pairs'] > 10 ** 6 and wa['streams'] >= 2
            assert wa['bins'] <= 64 * wa['streams']      # every stream in the warp shape
            assert np.array_equal(a, a2)        # fixed summation order: bit-identical repeats
            fa, dfa = unpack_terms(engine.desc, k, a)
            fb, dfb = unpack_terms(engine.desc, k, b)
            ref = {'k/%s' % label: val for label, val in iter_sum_entries(fb, dfb)}
            assert_sums_close(fa, dfa, ref, 'k', rtol=1e-11, atol=1e-10)
    finally:
        engine.use_lag_quadrature(True)
        engine.use_weight_cache(True)


def test_long_memory_kernel_pairs_with_every_earlier_event(engine):
    """An exponential kernel whose reach exceeds the observation window: every target pairs
    with every earlier event (2 x 10^8 pairs per stream, thousands of sort chunks per tile, the
    whole path near T_f in units of the kernel)."""
    from aslsd_b200.engine import unpack_terms
    rng = np.random.default_rng(31)
    T_f = 1000.
    times = [np.sort(rng.uniform(0, T_f, size=20000))]
    model = {'d': 1, 'kernels': [[[{'type': 'exp'}, {'type': 'exp'}]]]}
    engine.set_events(times, T_f)
    engine.set_model(model)
    x = np.array([0.5, 0.3, 0.002, 0.2, 1.5])
    try:
        engine.use_lag_quadrature(True)
        a = engine.exact_sums_packed(0, x)
        engine.use_lag_quadrature(False)
        b = engine.exact_sums_packed(0, x)
    finally:
        engine.use_lag_quadrature(True)
    fa, dfa = unpack_terms(engine.desc, 0, a)
    fb, dfb = unpack_terms(engine.desc, 0, b)
    ref = {'k/%s' % label: val for label, val in iter_sum_entries(fb, dfb)}
    assert_sums_close(fa, dfa, ref, 'k', rtol=1e-11, atol=1e-9)


def test_weight_cache_reuses_the_quadrature(engine):
    """With the weight cache on, a second evaluation on the same events and (beta, delta)
    skips the pass over the event pairs and gives the same sums; a new omega is honoured;
    new events invalidate the weights."""
    name = 'sbf2d3'
    model = CASES[name][0]
    rng = np.random.default_rng(8)
    times = _simulate_poisson_cluster(rng, 2, 5000., 1.2)
    engine.set_events(times, 5000.)
    engine.set_model(model)
    xs = case_x(name)
    try:
        engine.use_weight_cache(False)
        ref0 = engine.exact_sums_packed(0, xs[0])
        assert engine.last_work_counts()['moment_pairs'] > 0
        ref0b = engine.exact_sums_packed(0, xs[0])
        assert engine.last_work_counts()['moment_pairs'] > 0
        np.testing.assert_array_equal(ref0, ref0b)      # deterministic
        engine.use_weight_cache(True)
        a = engine.exact_sums_packed(0, xs[0])
        assert engine.last_work_counts()['moment_pairs'] > 0
        b = engine.exact_sums_packed(0, xs[0])
        assert engine.last_work_counts()['moment_pairs'] == 0
        np.testing.assert_array_equal(a, ref0)
        np.testing.assert_array_equal(b, ref0)
        x2 = xs[0].copy()
        x2[1:] *= 1.3
        c = engine.exact_sums_packed(0, x2)
        assert engine.last_work_counts()['moment_pairs'] == 0
        engine.use_weight_cache(False)
        np.testing.assert_array_equal(c, engine.exact_sums_packed(0, x2))
        engine.use_weight_cache(True)
        engine.exact_sums_packed(0, xs[0])
        times2 = _simulate_poisson_cluster(rng, 2, 5000., 1.2)
        engine.set_events(times2, 5000.)
        d2 = engine.exact_sums_packed(0, xs[0])
        assert engine.last_work_counts()['moment_pairs'] > 0
        assert abs(d2[0] - ref0[0]) > 0
    finally:
        engine.use_weight_cache(True)


def test_linearity_in_omega_for_sbf(engine):
    """Size-independent property: with (beta, delta) fixed every sum is a
    polynomial of degree <= 2 in omega; phi sums are linear."""
    name = 'sbf1d10'
    model = CASES[name][0]
    rng = np.random.default_rng(3)
    times = _simulate_poisson_cluster(rng, 1, 20000., 1.0)
    engine.set_events(times, 20000.)
    engine.set_model(model)
    x = case_x(name)[0]
    f1, _ = engine.exact_sums(0, x)
    x2 = x.copy()
    x2[1:] *= 2.0
    f2, _ = engine.exact_sums(0, x2)
    assert abs(f2['phi_0'] - 2 * f1['phi_0']) <= 1e-12 * abs(f1['phi_0'])
    assert abs(f2['upsilon_0_0'] - 4 * f1['upsilon_0_0']) <= 1e-12 * abs(f1['upsilon_0_0'])
    assert abs(f2['upsilonzero_0'] - 4 * f1['upsilonzero_0']) <= 1e-12 * abs(f1['upsilonzero_0'])


def test_errors(engine):
    import aslsd_b200
    g = load_golden('exact_exp1d')
    times, T_f = golden_times(g)
    engine.set_events(times, T_f)
    # kernels of different basis structure: the reference's late-bound upsilon closures would
    # read one kernel's parameters with another one's functional
    engine.set_events([times[0], times[0][::2]], T_f)
    with pytest.raises(NotImplementedError):
        engine.set_model({'d': 2, 'kernels': [[[{'type': 'exp'}], [{'type': 'exp'}]],
                                              [[{'type': 'exp'}], [{'type': 'gauss'}]]]})
    engine.set_events(times, T_f)
    engine.set_model(CASES['exp1d'][0])
    with pytest.raises(ValueError):
        engine.exact_sums(0, [1.0, 0.2])
    with pytest.raises(aslsd_b200.EngineError):
        engine.exact_sums(0, [1.0, 0.2, -1.0])


@pytest.fixture(scope='module')
def c5():
    """BASELINE.json's full-size configuration: 10-D SBF-Gaussian(10), ~10^7 events."""
    from aslsd_b200 import simulate
    return simulate.config_c5(d=10, n_basis=10, T_f=1.1e6, seed=1234)


def test_full_size_properties_c5(engine, c5):
    """Size-independent properties at the full benchmark size (the oracle cannot run there):
    time-block shards add up to the full sums; phi sums are linear and upsilon sums quadratic in
    omega; the lag quadrature agrees with the per-item pair kernels on every term of a
    sub-problem."""
    from aslsd_b200.distributed import time_block_ranges
    from aslsd_b200.engine import unpack_terms
    model, times, T_f, xs = c5
    d = model['d']
    engine.set_events(times, T_f)
    engine.set_model(model)
    ks = list(range(d))
    full = engine.exact_sums_batch(ks, xs)
    # (a) three time blocks add up
    acc = [np.zeros_like(f) for f in full]
    for block in time_block_ranges(times, 3):
        for a, part in zip(acc, engine.exact_sums_batch(ks, xs, ranges=block)):
            a += part
    for a, f in zip(acc, full):
        np.testing.assert_allclose(a, f, rtol=1e-11, atol=1e-9)
    # (b) scaling in omega (beta, delta are fixed in this model: every kernel parameter is an omega)
    k = 3
    f1, _ = unpack_terms(engine.desc, k, full[k])
    x2 = xs[k].copy()
    x2[1:] *= 2.0
    f2, _ = engine.exact_sums(k, x2)
    for i in range(d):
        assert abs(f2['phi_%d' % i] - 2 * f1['phi_%d' % i]) <= 1e-11 * abs(f1['phi_%d' % i])
        assert abs(f2['upsilonzero_%d' % i] - 4 * f1['upsilonzero_%d' % i]) <= 1e-11 * abs(f1['upsilonzero_%d' % i])
        for j in range(d):
            name = 'upsilon_%d_%d' % (i, j)
            assert abs(f2[name] - 4 * f1[name]) <= 1e-11 * abs(f1[name])
    # (c) quadrature against the per-item pair kernels, one sub-problem, every term
    try:
        engine.use_lag_quadrature(False)
        slow = engine.exact_sums_packed(k, xs[k])
    finally:
        engine.use_lag_quadrature(True)
    fa, dfa = unpack_terms(engine.desc, k, full[k])
    fb, dfb = unpack_terms(engine.desc, k, slow)
    ref = {'k/%s' % label: val for label, val in iter_sum_entries(fb, dfb)}
    assert_sums_close(fa, dfa, ref, 'k', rtol=1e-11, atol=1e-9)


def _config(name):
    """Configurations C2-C4 of BASELINE.json / SURVEY 8d at their named sizes."""
    from aslsd_b200 import simulate
    exp = lambda: [{'type': 'exp'}]
    if name == 'C2':      # 2-D exponential, ~10^5 events (fit_exp2d.py:34-51)
        model = {'d': 2, 'kernels': [[exp(), exp()], [exp(), exp()]]}
        mu, kp, T_f = [1.5, 1.], [[[.2, 1.], [.6, 1.5]], [[.7, 2.], [.1, 1.3]]], 7.7e3
    elif name == 'C3':    # 1-D Gaussian, ~10^6 events (fit_gauss1d.py:31-34)
        model = {'d': 1, 'kernels': [[[{'type': 'gauss'}]]]}
        mu, kp, T_f = [1.5], [[[.5, .5, 3.]]], 3.33e5
    else:                 # C4: 1-D SBF-Gaussian(10), ~10^6 events (fit_sbgauss1d10r.py:60-63)
        model = simulate.sbf_gauss_model(1, 10)
        mu, kp, T_f = [1.0], [[simulate.random_mixture_weights(10, coeff=0.5)]], 5e5
    times = simulate.simulate_mhp(model, mu, kp, T_f, seed=1234)
    d = model['d']
    rng = np.random.default_rng(12)
    xs = []
    for k in range(d):
        x = np.concatenate([[mu[k]]] + [np.asarray(kp[k][i], float) for i in range(d)])
        xs.append(x * (1 + 0.1 * rng.uniform(-1, 1, size=len(x))))
    return model, times, T_f, xs


@pytest.mark.parametrize('name', ['C2', 'C3', 'C4'])
def test_named_configurations_at_full_size(engine, name):
    """BASELINE.json's other configurations at their named sizes (10^5 - 10^6 events), where the
    O(N^2) oracle cannot run: the lag quadrature against the per-item pair kernels on every term
    and derivative, time-block shards adding up, and (C2: "exact vs stratified LSE gradient") the
    stratified estimate scattering around the exact gradient."""
    import aslsd_b200
    from aslsd_b200.distributed import time_block_ranges
    from aslsd_b200.engine import unpack_terms
    model, times, T_f, xs = _config(name)
    d = model['d']
    assert sum(len(t) for t in times) > (8e4 if name == 'C2' else 8e5)
    engine.set_events(times, T_f)
    engine.set_model(model)
    for k in range(d):
        fast = engine.exact_sums_packed(k, xs[k])
        assert engine.last_work_counts()['moment_pairs'] > 0
        try:
            engine.use_lag_quadrature(False)
            slow = engine.exact_sums_packed(k, xs[k])
        finally:
            engine.use_lag_quadrature(True)
        fa, dfa = unpack_terms(engine.desc, k, fast)
        fb, dfb = unpack_terms(engine.desc, k, slow)
        ref = {'k/%s' % label: val for label, val in iter_sum_entries(fb, dfb)}
        assert_sums_close(fa, dfa, ref, 'k', rtol=1e-11, atol=1e-8)
        acc = np.zeros_like(fast)
        for block in time_block_ranges(times, 4):
            acc += engine.exact_sums_packed(k, xs[k], ranges=block)
        np.testing.assert_allclose(acc, fast, rtol=1e-11, atol=1e-8)
    if name == 'C2':
        path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
        mirror = aslsd_b200.MHP(model)
        exact = aslsd_b200.MHPExactEstimB200(engine=engine, reference_recombine=False)
        exact.initialize(0, 2, mirror, path)
        g_exact = exact.lse_k_grad_estimate(xs[0])
        rng = np.random.default_rng(2)
        strat = aslsd_b200.MHPStratEstimB200(engine=engine, reference_adaptive_grad=False)
        strat.initialize(0, 40, mirror, path)
        g = np.mean([strat.lse_k_grad_estimate(xs[0], rng=rng, grad_alloc=False) for _ in range(40)], axis=0)
        assert np.linalg.norm(g - g_exact) <= 0.25 * np.linalg.norm(g_exact)


@pytest.fixture(scope='module')
def c5_short():
    """The benchmark's own model (10 event types, every kernel a sum of 10 Gaussians with fixed
    shapes: 1 111 packed sums per sub-problem) on a path short enough for the oracle, with the
    oracle's sums of every sub-problem (one host process each)."""
    from aslsd_b200 import simulate
    from helpers import oracle_sums_parallel
    model, times, T_f, xs = simulate.config_c5(d=10, n_basis=10, T_f=700., seed=1234)
    ref = oracle_sums_parallel(model, list(range(10)), xs, times, T_f)
    return model, times, T_f, xs, ref


def test_c5_model_matches_oracle_every_k(engine, c5_short):
    """Device against the oracle on the 10-type benchmark model, every sub-problem, every term and
    derivative at 1e-10 (mhp_exact_estimator.py:611-707): the single-k call, the batched call the
    benchmark times (rows shared between the sub-problems) and time-block shards."""
    from aslsd_b200.distributed import time_block_ranges
    from aslsd_b200.engine import unpack_terms
    model, times, T_f, xs, ref = c5_short
    d = model['d']
    engine.set_events(times, T_f)
    engine.set_model(model)
    ks = list(range(d))
    batch = engine.exact_sums_batch(ks, xs)
    work = engine.last_work_counts()
    assert work['moment_pairs'] > 0 and work['rows'] == 11440    # the quadrature path, C5's row table
    # the second call on the same row plan assembles on the device (linear map of the raw rows,
    # exact.cu assemble_kernel) and reads back the packed sums only; the first one did it on the host
    again = engine.exact_sums_batch(ks, xs)
    for a, b in zip(batch, again):
        np.testing.assert_allclose(b, a, rtol=1e-13, atol=1e-13 * np.abs(a).max())
    import torch
    dev = torch.zeros(engine.exact_packed_len_batch(ks), dtype=torch.float64, device='cuda:0')
    engine.exact_packed_device_batch(ks, xs, dev.data_ptr())
    engine.synchronize()
    for a, b in zip(batch, engine.split_packed(ks, dev.cpu().numpy())):
        np.testing.assert_allclose(b, a, rtol=1e-13, atol=1e-13 * np.abs(a).max())
    try:
        engine.use_device_assembly(False)
        host = engine.exact_sums_batch(ks, xs)
    finally:
        engine.use_device_assembly(True)
    for a, b in zip(batch, host):
        np.testing.assert_array_equal(a, b)
    blocks = time_block_ranges(times, 2)
    for k in ks:
        f, df = unpack_terms(engine.desc, k, again[k])
        assert_sums_close(f, df, ref[k][0], 'k', rtol=RTOL)
        f, df = unpack_terms(engine.desc, k, batch[k])
        assert_sums_close(f, df, ref[k][0], 'k', rtol=RTOL)
        f, df = engine.exact_sums(k, xs[k])
        assert_sums_close(f, df, ref[k][0], 'k', rtol=RTOL)
    for k in (0, 7):
        acc = sum(engine.exact_sums_packed(k, xs[k], ranges=b) for b in blocks)
        f, df = unpack_terms(engine.desc, k, acc)
        assert_sums_close(f, df, ref[k][0], 'k', rtol=RTOL, atol=1e-13)
    # the per-item kernels on the same model (no quadrature)
    try:
        engine.use_lag_quadrature(False)
        f, df = engine.exact_sums(4, xs[4])
    finally:
        engine.use_lag_quadrature(True)
    assert_sums_close(f, df, ref[4][0], 'k', rtol=RTOL)


def test_c5_model_slice_matches_oracle(engine, c5_short):
    """The parity gate of bench.py: targets of a time slice of sub-problem 0."""
    from helpers import oracle_sums_parallel
    from aslsd_b200.engine import unpack_terms
    model, times, T_f, xs, _ = c5_short
    engine.set_events(times, T_f)
    engine.set_model(model)
    ranges = [(int(np.searchsorted(t, 100.)), int(np.searchsorted(t, 450.))) for t in times]
    (ref, n_pairs), = oracle_sums_parallel(model, [0], [xs[0]], times, T_f, ranges=ranges, n_proc=1)
    packed, counts = engine.exact_sums_packed(0, xs[0], ranges=ranges, count_pairs=True)
    f, df = unpack_terms(engine.desc, 0, packed)
    assert_sums_close(f, df, ref, 'k', rtol=RTOL)
    for name, n in n_pairs.items():
        assert counts[name] == n, name
