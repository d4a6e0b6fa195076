"""Three (and four) event types on the device: the sub-problem index k differs from both i and j
in most upsilon_i_j terms, the rule in the 10-dimensional configuration C5 and impossible for
d <= 2.  Golden vectors of the literal reference (exact_exp3d / exact_sbf3d2 / exact_gauss3d,
tests/golden/make_golden.py 3d), the windowed oracle at ~10^4 events per type, sampled sums, the
batched call and the time-block shards.  Tolerance 1e-10 as everywhere.

(``python tests/test_gpu_three_types.py`` dry-runs the bodies that only need the engine interface
on the oracle test double.)"""
import numpy as np
import pytest

from cases import CASES, CASES_3D, case_x
from helpers import RTOL, iter_sum_entries, assert_sums_close
from oracle import exact as OE

import test_gpu_exact as TE
import test_gpu_stratified as TS

pytestmark = pytest.mark.gpu

NAMES = sorted(CASES_3D)


@pytest.fixture(scope='module')
def engine():
    import aslsd_b200
    return aslsd_b200.Engine(0)


@pytest.mark.parametrize('name', NAMES)
def test_exact_sums_match_reference_golden(engine, name):
    TE.test_exact_sums_match_reference_golden(engine, name)


@pytest.mark.parametrize('name', NAMES)
def test_lag_quadrature_agrees_with_item_kernels(engine, name):
    TE.test_lag_quadrature_agrees_with_item_kernels(engine, name)


def _four_types():
    """4-D model, every kernel an exponential plus a Gaussian basis (the kernel matrix has to be
    homogeneous: upstream builds every upsilon closure from the last kernel, mhp.py:527-536, SURVEY
    quirk 4), with different parameters in every cell."""
    from cases import square
    model = square(4, lambda: [{'type': 'exp'}, {'type': 'gauss'}])
    rng = np.random.default_rng(5)
    xs = []
    for k in range(4):
        x = [rng.uniform(.4, .9)]
        for i in range(4):
            x += [rng.uniform(.03, .12), rng.uniform(.8, 2.), rng.uniform(.03, .12), rng.uniform(.4, .8),
                  rng.uniform(.5, 2.)]
        xs.append(np.array(x))
    return model, xs


@pytest.mark.parametrize('name', NAMES + ['four_types'])
def test_exact_sums_match_windowed_oracle_medium(engine, name):
    if name == 'four_types':
        model, xs = _four_types()
        T_f, rate = 600., 2.0     # mixed pairs: lags kept short of the reference's exp overflow
    else:
        model, xs = CASES[name][0], case_x(name)
        T_f, rate = 2500., 1.5
    d = model['d']
    rng = np.random.default_rng(7)
    times = [np.sort(rng.uniform(0, T_f, size=rng.poisson(rate * (1 + 0.3 * i) * T_f))) for i in range(d)]
    engine.set_events(times, T_f)
    engine.set_model(model)
    for k in range(d):
        sum_f, sum_diff_f = engine.exact_sums(k, xs[k])
        o_f, o_df, _ = OE.exact_sums_windowed(model, k, xs[k], times, T_f)
        ref = {'k/%s' % label: val for label, val in iter_sum_entries(o_f, o_df)}
        assert_sums_close(sum_f, sum_diff_f, ref, 'k', rtol=RTOL)
    # the batched call shares rows between the sub-problems: same packed sums as one k at a time
    ks = list(range(d))
    batch = engine.exact_sums_batch(ks, xs)
    for k in ks:
        # (the quadrature bins of a batch differ from those of a single k: test_gpu_exact.py)
        np.testing.assert_allclose(batch[k], engine.exact_sums_packed(k, xs[k]), rtol=1e-10, atol=1e-9)
    # time-block shards add up
    from aslsd_b200.distributed import time_block_ranges
    acc = np.zeros_like(batch[1])
    for block in time_block_ranges(times, 3):
        acc += engine.exact_sums_packed(1, xs[1], ranges=block)
    np.testing.assert_allclose(acc, batch[1], rtol=1e-11, atol=1e-9)


@pytest.mark.parametrize('name', NAMES)
@pytest.mark.parametrize('grad_alloc', [False, True])
def test_sampled_sums_match_oracle(engine, name, grad_alloc):
    TS.test_sampled_sums_match_oracle(engine, name, grad_alloc)


@pytest.mark.parametrize('name', ['exp3d', 'sbf3d2'])
def test_estimators_on_three_types(engine, name):
    """Exact and stratified estimators through the mirror model: device against the oracle test
    double on the same draws."""
    import aslsd_b200
    from oracle_engine import OracleEngine
    model = CASES[name][0]
    rng = np.random.default_rng(3)
    T_f = 300.
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (500, 420, 610)]
    path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
    mirror = aslsd_b200.MHP(model)
    kw = dict(n_exact_single=50, n_samples_adaptive_single=60, nonadaptive_sample_size_single=80,
              n_samples_adaptive_double=200, nonadaptive_sample_size_double=90)
    for k in range(3):
        x = case_x(name)[k]
        dev = aslsd_b200.MHPExactEstimB200(engine=engine, reference_recombine=False)
        dev.initialize(k, 2, mirror, path)
        cpu = aslsd_b200.MHPExactEstimB200(engine=OracleEngine(), reference_recombine=False)
        cpu.initialize(k, 2, mirror, path)
        g_dev, g_cpu = dev.lse_k_grad_estimate(x), cpu.lse_k_grad_estimate(x)
        # the gradient is a signed combination of the sums (their own parity: the tests above)
        np.testing.assert_allclose(g_dev, g_cpu, rtol=1e-9, atol=1e-8 * np.abs(g_cpu).max())
        dev = aslsd_b200.MHPStratEstimB200(engine=engine)
        dev.initialize(k, 3, mirror, path, **kw)
        cpu = aslsd_b200.MHPStratEstimB200(engine=OracleEngine())
        cpu.initialize(k, 3, mirror, path, **kw)
        # first iteration: default allocations on both sides, hence the same draws; the following
        # ones depend on the rounding residue of near-constant strata (test_gpu_stratified.py)
        rng_a, rng_b = np.random.default_rng(9), np.random.default_rng(9)
        g_dev = dev.lse_k_grad_estimate(x, rng=rng_a, grad_alloc=False)
        g_cpu = cpu.lse_k_grad_estimate(x, rng=rng_b, grad_alloc=False)
        assert rng_a.bit_generator.state == rng_b.bit_generator.state
        np.testing.assert_allclose(g_dev, g_cpu, rtol=1e-9, atol=1e-9 * np.abs(g_cpu).max())
        for _ in range(2):
            assert np.all(np.isfinite(dev.lse_k_grad_estimate(x, rng=rng_a, grad_alloc=False)))


@pytest.mark.parametrize('name', ['exp3d', 'sbf3d2'])
def test_intensity_and_residuals_on_three_types(engine, name):
    """aslsd_point_sums with three event types against the oracle (mhp.py:669-690,
    goodness_of_fit.py:134-236)."""
    import aslsd_b200
    from aslsd_b200 import residuals as R
    from oracle import residuals as OR
    model, mu, kp, _ = CASES[name]
    rng = np.random.default_rng(12)
    T_f = 600.
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (700, 520, 610)]
    path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
    intensity = R.get_intensity_at_jumps(model, path, mu, kp, engine=engine)
    want = OR.intensity_at_jumps(model, mu, kp, times)
    for i in range(3):
        np.testing.assert_allclose(intensity[i], want[i], rtol=RTOL, atol=1e-13)
        got = R.get_residuals_k(model, path, i, mu, kp, engine=engine)
        np.testing.assert_allclose(got, OR.residuals_k(model, i, mu, kp, times), rtol=1e-9, atol=1e-9)


def test_marked_model_matches_reference_golden(engine):
    """mtlh_exp3d: golden vector of the reference's MTLHExactEstim with three event types."""
    import test_gpu_mtlh as TM
    TM.test_mtlh_exact_matches_reference_golden(engine, 'mtlh_exp3d')


if __name__ == '__main__':
    # dry run of the test bodies that only need the engine interface, on the oracle test double
    from oracle_engine import OracleEngine
    for n in ['exp3d', 'sbf3d2']:
        test_estimators_on_three_types(OracleEngine(), n)
        test_intensity_and_residuals_on_three_types(OracleEngine(), n)
    print('dry run ok')
