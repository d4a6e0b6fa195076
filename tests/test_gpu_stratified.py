"""Parity of the sampled-lag kernel and the fused allocation update (CUDA,
through the C ABI) against the stratified oracle and against golden traces of
the reference's MHPStratEstim.  Tolerance: relative 1e-10 (fp64)."""
import numpy as np
import pytest

from cases import CASES, case_x
from helpers import RTOL, replay_strat_golden
from oracle import stratified as OS

pytestmark = pytest.mark.gpu

ABS_FLOOR = 1e-290

NAMES = ['strat_exp1d', 'strat_exp2d', 'strat_exp2d_gradalloc', 'strat_gauss1d', 'strat_sbf2d3', 'strat_mix1d']


@pytest.fixture(scope='module')
def engine():
    import aslsd_b200
    return aslsd_b200.Engine(0)


@pytest.mark.parametrize('name', NAMES)
def test_replay_reference_trace_on_device(engine, name):
    """Samples forced to the reference's: every estimate within 1e-10, allocations of
    non-degenerate groups within 1e-8, RNG stream aligned to the end."""
    worst = replay_strat_golden(name, engine, mode='forced')
    assert worst <= RTOL


@pytest.mark.parametrize('name', ['strat_exp1d', 'strat_exp2d', 'strat_sbf2d3', 'strat_exp2d_long',
                                  'strat_gauss1d_long', 'strat_sbf2d3_long'])
def test_replay_reference_trace_strict(engine, name):
    """Nothing but the evaluation point is forced: every integer sample count and the whole RNG
    stream are those of the reference, over up to 120 iterations.  The samples of the
    saturated single-index strata carry the value of the model's own closures
    (aslsd_b200/saturation.py) and the statistic kernels follow numpy's summation order, so even
    the rounding-residue allocations of those strata are the reference's; exponential, Gaussian
    and sum-of-Gaussians kernels, grad_alloc=False (the fit default)."""
    replay_strat_golden(name, engine, mode='strict', free_running=(name == 'strat_exp1d'))


def test_replay_reference_trace_of_the_benchmark_model(engine):
    """The benchmark's own model (config C5: ten event types, SBF-Gaussian(10)): the reference's
    trace of three ADAM iterations of every sub-problem (tests/golden/strat_c5model.npz, 121 terms
    per call), nothing forced but the evaluation points - per-term sums, gradients, every adaptive
    allocation and the final generator state are the reference's (profiles/r2_zj_replay_c5model_device.log:
    8e-16, one second; the CPU suite replays the same trace over the oracle double)."""
    worst = replay_strat_golden('strat_c5model', engine, mode='strict')
    assert worst <= RTOL


def _random_batch(rng, times, k, d, n_strata, n_per, kinds=None):
    """Random (m, n) samples for every term kind, far lags included."""
    terms, offsets, ms, ns = [], [0], [], []
    if kinds is None:
        kinds = [(0, 0, 0), (1, d - 1, d - 1), (2, 0, 0), (2, d - 1, d - 1), (3, 0, 0), (3, 0, d - 1), (3, d - 1, 0)]
    for kind, i, j in kinds:
        tgt, src = (k, i) if kind == 2 else (i, j)
        terms.append((kind, i, j, n_strata))
        for s in range(n_strata):
            n_s = int(rng.integers(1, n_per))
            m = rng.integers(1, len(times[tgt]), size=n_s)
            if kind < 2:
                n = np.zeros(n_s, dtype=np.int64)
            else:
                hi = np.searchsorted(times[src], times[tgt][m], side='left')
                hi = np.maximum(hi, 1)
                lag = rng.integers(1, 40 if s % 2 == 0 else 2000, size=n_s)
                n = np.maximum(hi - lag, 0)
            ms.append(m)
            ns.append(n)
            offsets.append(offsets[-1] + n_s)
    return terms, np.array(offsets), np.concatenate(ms), np.concatenate(ns)


@pytest.mark.parametrize('name', ['exp2d', 'gauss2d', 'sbf2d3', 'exp1d2b', 'gauss1d2b', 'sbf1d10', 'mix1d', 'mix2d'])
@pytest.mark.parametrize('grad_alloc', [False, True])
def test_sampled_sums_match_oracle(engine, name, grad_alloc):
    model = CASES[name][0]
    d = model['d']
    rng = np.random.default_rng(11)
    # mixed pairs: the reference's exp x gauss value is exp(beta_1 s + ...) x (difference of cdfs),
    # inf x 0 = nan once beta_1 s passes ~709 (basis_kernel_exponential.py:30-39); keep the lags
    # of this comparison below that
    T_f = 300. if name.startswith('mix') else 3000.
    times = [np.sort(rng.uniform(0, T_f, size=4000 + 500 * i)) for i in range(d)]
    # a few targets close to T_f exercise the non-interior branches
    for i in range(d):
        times[i][-5:] = T_f - np.array([2.0, 1.0, 0.3, 0.05, 0.001]) * (1 + i)
        times[i].sort()
    engine.set_events(times, T_f)
    engine.set_model(model)
    for k in range(d):
        _check_sampled_against_oracle(engine, model, times, T_f, k, case_x(name)[k], rng, grad_alloc)


def _check_sampled_against_oracle(engine, model, times, T_f, k, x, rng, grad_alloc, kinds=None):
    d = model['d']
    terms, offsets, ix_m, ix_n = _random_batch(rng, times, k, d, n_strata=4, n_per=300, kinds=kinds)
    n_strata = len(offsets) - 1
    sizes = np.where(np.arange(n_strata) % 4 < 2, rng.integers(2, 10 ** 6, size=n_strata), 0).astype(float)
    dflt = rng.uniform(size=n_strata)
    sums, out_off, mean, std, rel = engine.sampled_sums(
        k, x, terms, offsets, ix_m, ix_n, strata_sizes=sizes, default_alloc=dflt,
        compute_diff_f_sum=True, grad_alloc=grad_alloc)
    s_ix = 0
    for kind, i, j, ns in terms:
        for s in range(ns):
            a, b = offsets[s_ix], offsets[s_ix + 1]
            vals = OS.term_values(model, k, x, times, T_f, kind, i, j, ix_m[a:b], ix_n[a:b])
            tot, o_mean, o_std = OS.stratum_statistics(vals, grad_alloc)
            got = sums[out_off[s_ix]:out_off[s_ix + 1]]
            scale = np.abs(vals).sum(axis=0)
            # absolute floor ("abs floor 1e-300"; here 1e-290): a Gaussian stratum
            # whose samples all lie in the far tail sums to a denormal (< 2.2e-308), where
            # exp of the device and of numpy differ in the few bits that are left
            err = np.abs(got - tot)
            ok = err <= RTOL * np.abs(tot) + 1e-13 * scale
            # ... and the floor is only ever what decides in that range
            assert np.all(scale[~ok] < ABS_FLOOR), (kind, i, j, s, err, scale)
            assert np.all(err <= RTOL * np.abs(tot) + 1e-13 * scale + ABS_FLOOR), (kind, i, j, s)
            assert abs(mean[s_ix] - o_mean) <= RTOL * abs(o_mean) + ABS_FLOOR
            assert abs(std[s_ix] - o_std) <= 1e-9 * abs(o_std) + 1e-14 * abs(o_mean) + ABS_FLOOR
            s_ix += 1
    # allocation groups: strata 0-1 of every term
    for q in range(len(terms)):
        a = 4 * q
        want = OS.update_allocation(sizes[a:a + 2], std[a:a + 2], dflt[a:a + 2])
        np.testing.assert_allclose(rel[a:a + 2], want, rtol=1e-12)
        np.testing.assert_array_equal(rel[a + 2:a + 4], dflt[a + 2:a + 4])


@pytest.mark.parametrize('grad_alloc', [False, True])
def test_sampled_sums_of_the_benchmark_model_match_oracle(engine, grad_alloc):
    """The benchmark's own model (ten event types, sums of ten Gaussians, configuration C5) on a
    short path: per-stratum sums, statistics and allocations of terms whose event types differ from
    the sub-problem's and from each other (the rule at d = 10), against the oracle."""
    from aslsd_b200 import simulate
    model, times, T_f, xs = simulate.config_c5(d=10, n_basis=10, T_f=700., seed=1234)
    engine.set_events(times, T_f)
    engine.set_model(model)
    rng = np.random.default_rng(17)
    for k, kinds in ((0, [(0, 3, 3), (1, 7, 7), (2, 5, 5), (2, 0, 0), (3, 4, 4), (3, 2, 8), (3, 8, 2), (3, 0, 6)]),
                     (6, [(0, 6, 6), (1, 1, 1), (2, 9, 9), (3, 6, 6), (3, 9, 3), (3, 3, 6)])):
        _check_sampled_against_oracle(engine, model, times, T_f, k, xs[k], rng, grad_alloc, kinds=kinds)


def test_update_allocation_entry_point(engine):
    sizes = np.array([10., 200., 3000., 2.])
    std = np.array([0.5, 0.1, 0.01, 3.])
    dflt = np.array([.25, .25, .25, .25])
    got = engine.update_allocation(sizes, std, dflt)
    np.testing.assert_allclose(got, OS.update_allocation(sizes, std, dflt), rtol=1e-14)
    # every std zero -> previous allocation kept (mhp_stratified_estimator.py:641-645)
    np.testing.assert_array_equal(engine.update_allocation(sizes, np.zeros(4), dflt), dflt)
    # a stratum of size 1 gives inf * 0 = nan upstream -> default allocation
    np.testing.assert_array_equal(engine.update_allocation(np.array([1., 5.]), np.array([0., 1.]),
                                                           np.array([.3, .7])), [.3, .7])


def test_sampled_sums_reject_bad_indices(engine):
    import aslsd_b200
    rng = np.random.default_rng(0)
    times = [np.sort(rng.uniform(0, 100, size=50))]
    engine.set_events(times, 100.)
    engine.set_model(CASES['exp1d'][0])
    x = case_x('exp1d')[0]
    with pytest.raises(aslsd_b200.EngineError):
        engine.sampled_sums(0, x, [(3, 0, 0, 1)], [0, 2], [3, 50], [1, 2])
    with pytest.raises(aslsd_b200.EngineError):
        engine.sampled_sums(0, x, [(2, 0, 0, 1)], [0, 2], [3, 4], [-1, 2])
    with pytest.raises(aslsd_b200.EngineError):
        engine.sampled_sums(0, x, [(7, 0, 0, 1)], [0, 1], [3], [1])


def test_mirror_fit_runs_on_device(engine):
    """End-to-end ASLSD fit loop (host mirror + CUDA), free running.  The first
    iterates reproduce the reference's; later ones only statistically: a last-bit
    difference in an iterate changes the rounding-residue allocations of the
    near-constant strata, hence the sample counts and the RNG stream."""
    import aslsd_b200
    from helpers import load_golden
    g = load_golden('strat_exp1d')
    times = [np.asarray(g['times/0'])]
    path = aslsd_b200.ProcessPath(times, float(g['T_f']), engine=engine)
    model = aslsd_b200.MHP(CASES['exp1d'][0])
    n_iter = int(g['n_iter'])
    x, log = model.fit(path, [g['k0/x'][0]], n_iter=n_iter, seed=int(g['seed']), engine=engine)
    np.testing.assert_allclose(log['param'][0][:3], g['k0/x'][:3], rtol=1e-9)
    np.testing.assert_allclose(log['grad'][0][:2], g['k0/grad'][:2], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(log['param'][0], g['k0/x'], rtol=0.05)
    assert model.fit_estim[0].t == n_iter


def test_allocation_statistic_is_bitwise_numpy(engine):
    """psi of an exponential kernel far from T_f is the same constant for every
    sample: the reference's std is then the rounding residue of numpy's pairwise
    sum, and the next integer allocations depend on it.  The device reproduces
    np.mean / np.std bit for bit on bit-identical values."""
    rng = np.random.default_rng(5)
    T_f = 5000.
    times = [np.sort(rng.uniform(0, T_f - 200., size=6000))]
    engine.set_events(times, T_f)
    engine.set_model(CASES['exp1d'][0])
    x = np.array([1.37, 0.31, 0.83])
    counts = [1, 7, 8, 9, 100, 128, 129, 345, 1000, 2999]
    offsets = np.concatenate([[0], np.cumsum(counts)])
    ix_m = rng.integers(0, 6000, size=offsets[-1])
    sums, out_off, mean, std, _ = engine.sampled_sums(0, x, [(0, 0, 0, len(counts))], offsets, ix_m,
                                                      np.zeros_like(ix_m), compute_diff_f_sum=True)
    for s, n in enumerate(counts):
        vals = OS.term_values(CASES['exp1d'][0], 0, x, times, T_f, 0, 0, 0,
                              ix_m[offsets[s]:offsets[s + 1]])[:, 0]
        assert np.all(vals == vals[0])           # degenerate stratum
        assert mean[s] == np.mean(vals), (s, n)
        assert std[s] == np.std(vals), (s, n)
    # generic values: still bitwise, given the device's own per-sample values
    T_f2 = 300.
    times2 = [np.sort(rng.uniform(0, T_f2, size=3000))]
    engine.set_events(times2, T_f2)
    ix = rng.integers(0, 3000, size=offsets[-1])
    sums, out_off, mean, std, _ = engine.sampled_sums(0, x, [(0, 0, 0, len(counts))], offsets, ix,
                                                      np.zeros_like(ix), compute_diff_f_sum=True)
    for s, n in enumerate(counts):
        vals = OS.term_values(CASES['exp1d'][0], 0, x, times2, T_f2, 0, 0, 0,
                              ix[offsets[s]:offsets[s + 1]])[:, 0]
        assert abs(mean[s] - np.mean(vals)) <= 1e-14 * abs(np.mean(vals))
        assert abs(std[s] - np.std(vals)) <= 1e-10 * np.std(vals) + 1e-16
