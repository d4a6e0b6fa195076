"""Device tests of the estimator / container surface: residuals of marked models,
ProcessPath.truncate, the per-term entry points of the estimators.  (``python
tests/test_gpu_surface.py`` dry-runs them on the oracle test double.)"""
import numpy as np
import pytest

from cases import CASES, case_x
from helpers import RTOL, load_golden
from oracle import time_ordering as OT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
    import aslsd_b200
    return aslsd_b200.Engine(0)


@pytest.mark.parametrize('name', ['mtlh_exp2d', 'mtlh_gauss1d', 'mtlh_sbf2d3'])
def test_marked_models_match_reference_golden(engine, name):
    """goodness_of_fit.get_residuals_mtlh: psi sums on the device, expected impacts on the host."""
    import aslsd_b200
    from cases import MTLH_CASES
    model = MTLH_CASES[name][0]
    g = load_golden('residuals_' + name)
    d = model['d']
    times = [g['times/%d' % i] for i in range(d)]
    kp = [[g['kernel_param/%d/%d' % (k, i)] for i in range(d)] for k in range(d)]
    mu = [g['mu/%d' % k] for k in range(d)]
    path = aslsd_b200.ProcessPath(times, float(g['T_f']), engine=engine)
    got = aslsd_b200.get_residuals_mtlh({'d': d, 'kernels': model['kernels']}, path, mu, kp,
                                        g['expected_impact'], engine=engine)
    for k in range(d):
        scale = np.max(np.abs(g['residuals/%d' % k]))
        np.testing.assert_allclose(got[k], g['residuals/%d' % k], rtol=RTOL, atol=1e-10 * scale)


def test_truncated_path_has_its_own_context_and_fresh_maps(engine):
    """ProcessPath.truncate (process_path.py:132-170): maps of the shorter path from a second
    device context, the full path stays resident in the first."""
    import aslsd_b200
    rng = np.random.default_rng(11)
    T_f = 500.
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (4000, 2500, 3100)]
    path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
    cut = path.truncate(187.3)
    assert cut.engine is not engine and cut.T_f == 187.3
    kept = [t[t < 187.3] for t in times]
    kappa, varpi = OT.kappa_varpi_closed_form(kept, 187.3)
    lags = OT.lag_sizes(kept, kappa, varpi)
    for i in range(3):
        np.testing.assert_array_equal(cut.list_times[i], kept[i])
        for j in range(3):
            np.testing.assert_array_equal(cut.kappa[i][j], kappa[i][j])
            np.testing.assert_array_equal(cut.varpi[i][j], varpi[i][j])
            np.testing.assert_array_equal(cut.lag_sizes[i][j], lags[i][j])
    # the first context still holds the full path
    g_kappa, _, _ = engine.index_maps()
    np.testing.assert_array_equal(g_kappa[0][1], path.kappa[0][1])


def test_per_term_entry_points_on_device(engine):
    """compute_*_sum one term at a time (mhp_exact_estimator.py:543-707,
    mhp_stratified_estimator.py:1081-1331): the device results equal the oracle test double's on the
    same draws, and the exact ones equal the batched call's."""
    import aslsd_b200
    from oracle_engine import OracleEngine
    name = 'sbf2d3'
    model = CASES[name][0]
    rng = np.random.default_rng(4)
    T_f = 400.
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (700, 500)]
    path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
    mirror = aslsd_b200.MHP(model)
    k = 1
    x = case_x(name)[k]
    lo0, hi0 = 1, 1 + 3
    x_k0, x_k1, mu_k = x[lo0:hi0], x[hi0:], x[0]

    def run(est, **kw):
        est.clear_estimates()
        est.compute_psi_sum(0, x_k0, mu_k, **kw)
        est.compute_phi_sum(1, x_k1, **kw)
        est.compute_cross_upsilon_sum(0, 1, x_k0, x_k1, **kw)
        est.compute_sim_upsilon_sum(1, x_k1, **kw)
        est.compute_upsilonzero_sum(0, x_k0, **kw)
    touched = ['psi_0', 'phi_1', 'upsilon_0_1', 'upsilon_1_1', 'upsilonzero_0']
    exact = aslsd_b200.MHPExactEstimB200(engine=engine)
    exact.initialize(k, 2, mirror, path)
    run(exact)
    per_term = {n: (exact.sum_f[n], exact.sum_diff_f[n]) for n in touched}
    exact.compute_objective(x, compute_f_sum=True, compute_diff_f_sum=True)
    for n in touched:
        np.testing.assert_allclose(per_term[n][0], exact.sum_f[n], rtol=1e-12)
        a, b = per_term[n][1], exact.sum_diff_f[n]
        for u, v in (zip(a[1:], b[1:]) if isinstance(a, list) else [(a, b)]):
            np.testing.assert_allclose(u, v, rtol=1e-11, atol=1e-12 * np.abs(v).max())
    kw = dict(n_exact_single=50, n_samples_adaptive_single=60, nonadaptive_sample_size_single=80,
              n_samples_adaptive_double=200, nonadaptive_sample_size_double=90)
    dev = aslsd_b200.MHPStratEstimB200(engine=engine)
    dev.initialize(k, 3, mirror, path, **kw)
    cpu = aslsd_b200.MHPStratEstimB200(engine=OracleEngine())
    cpu.initialize(k, 3, mirror, path, **kw)
    # one pass: the first allocations are the defaults on both sides, so the draws are the same
    # (later passes use allocations updated from the measured deviations, whose rounding residue
    # on near-constant strata differs between the device and numpy: test_replay_reference_trace_*)
    rng_a, rng_b = np.random.default_rng(21), np.random.default_rng(21)
    run(dev, rng=rng_a, grad_alloc=False, compute_f_sum=True)
    run(cpu, rng=rng_b, grad_alloc=False, compute_f_sum=True)
    assert rng_a.bit_generator.state == rng_b.bit_generator.state
    for n in touched:
        np.testing.assert_allclose(dev.sum_f[n], cpu.sum_f[n], rtol=RTOL)
        a, b = dev.sum_diff_f[n], cpu.sum_diff_f[n]
        for u, v in (zip(a[1:], b[1:]) if isinstance(a, list) else [(a, b)]):
            np.testing.assert_allclose(u, v, rtol=RTOL, atol=1e-10 * max(np.abs(v).max(), 1e-300))
    run(dev, rng=rng_a, grad_alloc=False, compute_f_sum=True)      # and a second pass runs
    assert all(np.isfinite(dev.sum_f[n]) for n in touched)


if __name__ == '__main__':
    from oracle_engine import OracleEngine
    for n in ['mtlh_exp2d', 'mtlh_gauss1d', 'mtlh_sbf2d3']:
        test_marked_models_match_reference_golden(OracleEngine(), n)
    test_truncated_path_has_its_own_context_and_fresh_maps(OracleEngine())
    test_per_term_entry_points_on_device(OracleEngine())
    print('dry run ok')
