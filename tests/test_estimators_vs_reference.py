"""Host logic of the B200 estimators against the LIVE reference (CPU, needs
/root/reference; skipped elsewhere).  The device is replaced by the oracle
test double (tests/oracle_engine.py), so what is checked here is everything
around the kernels: RNG consumption order, integer sample allocations, EMA
allocation state, result attributes and their shapes, recombination, logs, and
the use of the estimators as plug-ins of the reference's own ``MHP.fit``.
"""
import copy

import numpy as np
import pytest

import ref_shim
from cases import CASES, case_x
from oracle_engine import OracleEngine

pytestmark = pytest.mark.reference


@pytest.fixture(scope='module')
def ref():
    return ref_shim.import_reference()


def _simulate(ref, name, T_f):
    model, mu, kp, _ = CASES[name]
    d = model['d']
    mhp = ref_shim.build_reference_model(ref, model)
    arr = np.empty((d, d), dtype=object)
    for a in range(d):
        for b in range(d):
            arr[a][b] = np.asarray(kp[a][b], dtype=float)
    path = mhp.simulate(T_f, mu=np.asarray(mu, float), kernel_param=arr, seed=1234, verbose=False)
    return model, mhp, path


def _flat(d):
    out = []
    for name in sorted(d):
        v = d[name]
        if isinstance(v, list):
            out += [np.ravel(v[1]), np.ravel(v[2])]
        else:
            out.append(np.ravel(np.asarray(v, dtype=float)))
    return np.concatenate(out)


STRAT_KW = dict(n_exact_single=50, n_samples_adaptive_single=60, nonadaptive_sample_size_single=80,
                n_samples_adaptive_double=200, nonadaptive_sample_size_double=90)


@pytest.mark.parametrize('name,T_f,grad_alloc', [('exp1d', 600., False), ('exp2d', 150., False),
                                                 ('exp2d', 150., True), ('sbf2d3', 300., False),
                                                 ('gauss1d', 300., True)])
def test_stratified_estimator_tracks_reference(ref, name, T_f, grad_alloc):
    import aslsd_b200
    model, mhp, path = _simulate(ref, name, T_f)
    d = model['d']
    xs = case_x(name)
    rng_ref = np.random.default_rng(77)
    rng_new = np.random.default_rng(77)
    engine = OracleEngine()
    for k in range(d):
        est_ref = ref.MHPStratEstim(is_log_lse=True, is_log_allocs=True)
        est_ref.initialize(k, 12, mhp, path, **STRAT_KW)
        est_new = aslsd_b200.MHPStratEstimB200(is_log_lse=True, is_log_allocs=True, engine=engine)
        est_new.initialize(k, 12, mhp, path, **STRAT_KW)
        # same stratification geometry
        for log_name in est_ref.list_log_names:
            a, b = est_ref.stratif[log_name], est_new.stratif[log_name]
            assert a.n_adaptive_strata == b.n_adaptive_strata
            assert a.n_nonadaptive_strata == b.n_nonadaptive_strata
            assert a.n_exact == b.n_exact
            if a.n_adaptive_strata:
                np.testing.assert_array_equal(np.asarray(a.adaptive_strata), np.asarray(b.adaptive_strata))
                np.testing.assert_array_equal(a.adaptive_strata_sizes, b.adaptive_strata_sizes)
                np.testing.assert_array_equal(a.n_samples_adaptive, b.n_samples_adaptive)
            if a.n_nonadaptive_strata:
                np.testing.assert_array_equal(np.asarray(a.nonadaptive_strata),
                                              np.asarray(b.nonadaptive_strata))
                np.testing.assert_array_equal(a.nonadaptive_abs_alloc, b.nonadaptive_abs_alloc)
        x = np.array(xs[k], dtype=object)
        for t in range(12):
            g_ref = est_ref.lse_k_grad_estimate(x, rng=rng_ref, grad_alloc=grad_alloc)
            g_new = est_new.lse_k_grad_estimate(x, rng=rng_new, grad_alloc=grad_alloc)
            # identical RNG streams <=> identical integer allocations at every step
            assert rng_ref.bit_generator.state == rng_new.bit_generator.state, (k, t)
            scale = np.abs(_flat(est_ref.sum_diff_f)).max() / path.T_f
            np.testing.assert_allclose(g_new, g_ref, rtol=1e-10, atol=1e-10 * scale)
            np.testing.assert_allclose(_flat(est_new.sum_diff_f), _flat(est_ref.sum_diff_f),
                                       rtol=1e-10, atol=1e-12 * scale * path.T_f)
            np.testing.assert_allclose(_flat(est_new.sum_f), _flat(est_ref.sum_f), rtol=1e-10)
            for log_name in est_ref.list_log_names:
                a, b = est_ref.stratif[log_name], est_new.stratif[log_name]
                if a.n_adaptive_strata:
                    np.testing.assert_allclose(b.adaptive_rel_alloc, a.adaptive_rel_alloc,
                                               rtol=1e-9, atol=1e-15)
                    np.testing.assert_allclose(b.guess_rel_alloc, a.guess_rel_alloc, rtol=1e-9,
                                               atol=1e-15)
                np.testing.assert_allclose(est_new.vecsum_f_adaptive[log_name],
                                           est_ref.vecsum_f_adaptive[log_name], rtol=1e-10,
                                           atol=1e-300)
                np.testing.assert_allclose(est_new.sum_f_exact[log_name],
                                           est_ref.sum_f_exact[log_name], rtol=1e-10)
            x = x * (1. + 0.01 * np.cos(t + np.arange(len(x))))   # move the evaluation point
        np.testing.assert_allclose(est_new.logged_lse, est_ref.logged_lse, rtol=1e-10)
        assert est_new.t == est_ref.t
        assert set(est_new.get_log()) == set(est_ref.get_log())


@pytest.mark.parametrize('name,T_f', [('exp2d', 60.), ('sbf2d3', 100.), ('gauss1d', 100.)])
def test_exact_estimator_matches_reference(ref, name, T_f):
    import aslsd_b200
    model, mhp, path = _simulate(ref, name, T_f)
    engine = OracleEngine()
    for k in range(model['d']):
        x = np.array(case_x(name)[k], dtype=object)
        est_ref = ref.MHPExactEstim(is_log_lse=True)
        est_ref.initialize(k, 3, mhp, path)
        est_new = aslsd_b200.MHPExactEstimB200(is_log_lse=True, engine=engine)
        est_new.initialize(k, 3, mhp, path)
        g_ref = est_ref.lse_k_grad_estimate(x)
        g_new = est_new.lse_k_grad_estimate(x)
        scale = np.abs(_flat(est_ref.sum_diff_f)).max() / path.T_f
        # the reference's own (overwriting) gradient assembly is reproduced by default
        np.testing.assert_allclose(g_new, g_ref, rtol=1e-10, atol=1e-10 * scale)
        np.testing.assert_allclose(_flat(est_new.sum_diff_f), _flat(est_ref.sum_diff_f), rtol=1e-10,
                                   atol=1e-12 * scale * path.T_f)
        fixed = aslsd_b200.MHPExactEstimB200(engine=engine, reference_recombine=False)
        fixed.initialize(k, 3, mhp, path)
        g_fix = fixed.lse_k_grad_estimate(x)
        est_ref.recombine_diff_lse_theta_kpr = est_ref.recombine_diff_lse_theta_kpr_kernel
        g_strat_loop = ref.MHPStratEstim.recombine_grad_lse_k(est_ref)
        np.testing.assert_allclose(g_fix, g_strat_loop, rtol=1e-10, atol=1e-10 * scale)
        np.testing.assert_allclose(est_new.lse_k_estimate(x), est_ref.lse_k_estimate(x), rtol=1e-10)
        np.testing.assert_allclose(est_new.logged_lse, est_ref.logged_lse, rtol=1e-10)


def test_plugs_into_reference_fit(ref):
    """The estimators are drop-ins of MHP.fit(estimators=...): a single instance is
    deep-copied per dimension (subclass of the reference ABC), or a list is used."""
    import aslsd_b200
    from aslsd.optimize.estimators.estimator import Estimator
    assert issubclass(aslsd_b200.MHPStratEstimB200, Estimator)
    model, mhp, path = _simulate(ref, 'exp2d', 120.)
    x_0 = [np.array(x, dtype=object) for x in case_x('exp2d')]
    kw = dict(x_0=copy.deepcopy(x_0), n_iter=8, seed=99, **STRAT_KW)
    mu_ref, ker_ref = mhp.fit(path, **kw)
    engine = OracleEngine()
    mhp2 = ref_shim.build_reference_model(ref, model)
    mu_new, ker_new = mhp2.fit(path, estimators=aslsd_b200.MHPStratEstimB200(engine=engine),
                               **dict(kw, x_0=copy.deepcopy(x_0)))
    np.testing.assert_allclose(np.asarray(mu_new, float), np.asarray(mu_ref, float), rtol=1e-8)
    for a in range(2):
        for b in range(2):
            np.testing.assert_allclose(np.asarray(ker_new[a][b], float),
                                       np.asarray(ker_ref[a][b], float), rtol=1e-8)
    assert engine.n_calls == 16   # one device call per lse_k_grad_estimate


def test_mirror_fit_matches_reference_fit(ref):
    """aslsd_b200.MHP.fit (mirror loop + ADAM) reproduces the reference's fit."""
    import aslsd_b200
    model, mhp, path = _simulate(ref, 'exp1d', 500.)
    x_0 = [np.array(x, dtype=object) for x in case_x('exp1d')]
    mu_ref, ker_ref = mhp.fit(path, x_0=copy.deepcopy(x_0), n_iter=15, seed=5, **STRAT_KW)
    mirror = aslsd_b200.MHP(model)
    x, log = mirror.fit(path, [np.asarray(v, float) for v in case_x('exp1d')], n_iter=15, seed=5,
                        engine=OracleEngine(), **STRAT_KW)
    np.testing.assert_allclose(x[0][0], float(mu_ref[0]), rtol=1e-8)
    np.testing.assert_allclose(x[0][1:], np.asarray(ker_ref[0][0], float), rtol=1e-8)


def test_fit_episodes_matches_reference(ref):
    """MHP.fit_episodes (mhp.py:1806-2008) with exact estimators: the mirror loop reproduces the
    reference's iterates, and the B200 estimators plug into the reference's own fit_episodes."""
    import aslsd_b200
    from aslsd.stats.events.episodes import Episodes
    model, mu, kp, _ = CASES['exp2d']
    d = model['d']
    mhp = ref_shim.build_reference_model(ref, model)
    arr = np.empty((d, d), dtype=object)
    for a in range(d):
        for b in range(d):
            arr[a][b] = np.asarray(kp[a][b], dtype=float)
    paths = [mhp.simulate(50., mu=np.asarray(mu, float), kernel_param=arr, seed=100 + q, verbose=False)
             for q in range(3)]
    episodes = Episodes(list_paths=paths, T_f=50.)
    x_0 = [np.array(x, dtype=object) for x in case_x('exp2d')]
    mu_ref, ker_ref = mhp.fit_episodes(episodes, x_0=copy.deepcopy(x_0), n_iter=5, exact_grad=True)
    # (1) the B200 estimators inside the reference's loop; one oracle engine per path
    engines = [OracleEngine() for _ in paths]
    ests = [[aslsd_b200.MHPExactEstimB200(engine=engines[q]) for q in range(3)] for _ in range(d)]
    mhp2 = ref_shim.build_reference_model(ref, model)
    mu_new, ker_new = mhp2.fit_episodes(episodes, x_0=copy.deepcopy(x_0), n_iter=5, estimators=ests)
    np.testing.assert_allclose(np.asarray(mu_new, float), np.asarray(mu_ref, float), rtol=1e-9)
    for a in range(d):
        for b in range(d):
            np.testing.assert_allclose(np.asarray(ker_new[a][b], float), np.asarray(ker_ref[a][b], float),
                                       rtol=1e-9)
    # (2) the mirror loop
    mirror = aslsd_b200.MHP(model)
    engines = [OracleEngine() for _ in paths]
    ests = [[aslsd_b200.MHPExactEstimB200(engine=engines[q]) for q in range(3)] for _ in range(d)]
    x, log = mirror.fit_episodes(aslsd_b200.Episodes(paths, T_f=50.),
                                 [np.asarray(v, float) for v in case_x('exp2d')], n_iter=5, estimators=ests)
    for k in range(d):
        np.testing.assert_allclose(x[k][0], float(mu_ref[k]), rtol=1e-9)
        pos = 1
        for i in range(d):
            ref_ki = np.asarray(ker_ref[k][i], float)
            np.testing.assert_allclose(x[k][pos:pos + len(ref_ki)], ref_ki, rtol=1e-9)
            pos += len(ref_ki)
        assert log['grad'][k].shape == (5, len(x[k]))
