"""Host logic of the B200 estimators against the LIVE reference (CPU, needs
/root/reference; skipped elsewhere).  The device is replaced by the oracle
test double (tests/oracle_engine.py), so what is checked here is everything
around the kernels: RNG consumption order, integer sample allocations, EMA
allocation state, result attributes and their shapes, recombination, logs, and
the use of the estimators as plug-ins of the reference's own ``MHP.fit``.
"""
import copy
import os

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
                                                 ('gauss1d', 300., True), ('sbf3d2', 120., False)])
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


@pytest.mark.parametrize('name,T_f', [('exp2d', 60.), ('sbf2d3', 100.), ('gauss1d', 100.), ('exp3d', 50.),
                                      ('gauss3d', 40.)])
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


def _assert_same(a, b, where=''):
    """Recursive comparison of nested logs: same container types and keys, integer arrays equal,
    floating-point values within 1e-10."""
    if isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a) == sorted(b), where
        for key in a:
            _assert_same(a[key], b[key], '%s[%r]' % (where, key))
    elif isinstance(a, (list, tuple)):
        assert isinstance(b, (list, tuple)) and len(a) == len(b), where
        for q, (u, v) in enumerate(zip(a, b)):
            _assert_same(u, v, '%s[%d]' % (where, q))
    elif a is None:
        assert b is None, where
    else:
        u, v = np.asarray(a), np.asarray(b)
        assert u.shape == v.shape, where
        if u.dtype.kind in 'iu':
            np.testing.assert_array_equal(v, u, err_msg=where)
        else:
            scale = np.abs(u.astype(float)).max() if u.size else 0.
            np.testing.assert_allclose(v.astype(float), u.astype(float), rtol=1e-10, atol=1e-10 * scale,
                                       err_msg=where)


ALL_LOGS = dict(is_log_lse=True, is_log_allocs=True, is_log_ixs=True, is_log_total_estimates=True,
                is_log_strata_estimates=True)


@pytest.mark.parametrize('name,T_f', [('exp2d', 150.), ('gauss1d', 300.)])
def test_stratified_estimator_logs_match_reference(ref, name, T_f, tmp_path):
    """Every optional log of the reference's estimator (sampled indices, allocations, total and
    per-stratum estimates: mhp_stratified_estimator.py:272-474) holds the same values in the
    same containers, and the estimator survives its own save / load."""
    import aslsd_b200
    model, mhp, path = _simulate(ref, name, T_f)
    for k in range(model['d']):
        est_ref = ref.MHPStratEstim(**ALL_LOGS)
        est_ref.initialize(k, 5, mhp, path, **STRAT_KW)
        est_new = aslsd_b200.MHPStratEstimB200(engine=OracleEngine(), **ALL_LOGS)
        est_new.initialize(k, 5, mhp, path, **STRAT_KW)
        rng_ref, rng_new = np.random.default_rng(3), np.random.default_rng(3)
        x = np.array(case_x(name)[k], dtype=object)
        for t in range(4):
            est_ref.lse_k_grad_estimate(x, rng=rng_ref, grad_alloc=False)
            est_new.lse_k_grad_estimate(x, rng=rng_new, grad_alloc=False)
            x = x * (1. + 0.01 * np.cos(t + np.arange(len(x))))
        _assert_same(est_ref.get_log(), est_new.get_log(), 'get_log()')
        log_attrs = [a for a in vars(est_ref) if a.endswith('_log')]
        assert len(log_attrs) == 14 and sorted(log_attrs) == sorted(a for a in vars(est_new) if a.endswith('_log'))
        for attr in log_attrs:
            _assert_same(getattr(est_ref, attr), getattr(est_new, attr), attr)
        table_ref, names_ref = est_ref.get_lse_k_components_log()
        table_new, names_new = est_new.get_lse_k_components_log()
        assert names_new == names_ref
        _assert_same(table_ref, table_new, 'lse_k components')
        # the objective alone (lse_k_estimate, :1391-1405), drawn from the same generator state
        np.testing.assert_allclose(est_new.lse_k_estimate(x, rng=rng_new, grad_alloc=False, count_iter=False),
                                   est_ref.lse_k_estimate(x, rng=rng_ref, grad_alloc=False, count_iter=False),
                                   rtol=1e-10)
        assert rng_ref.bit_generator.state == rng_new.bit_generator.state
        # save / load in the reference's file layout (:1509-1689): own round trip, and files written
        # by either implementation load into the other
        file_new, file_ref = str(tmp_path / ('new_%d.pickle' % k)), str(tmp_path / ('ref_%d.pickle' % k))
        est_new.save(file_new)
        est_ref.save(file_ref)
        assert sorted(f.replace('new_', '') for f in os.listdir(tmp_path) if f.startswith('new_%d' % k)) == \
            sorted(f.replace('ref_', '') for f in os.listdir(tmp_path) if f.startswith('ref_%d' % k))
        for source, make in ((file_new, lambda: aslsd_b200.MHPStratEstimB200(engine=OracleEngine())),
                             (file_ref, lambda: aslsd_b200.MHPStratEstimB200(engine=OracleEngine())),
                             (file_new, ref.MHPStratEstim)):
            back = make()
            back.load(source, mhp, path)
            assert back.t == est_ref.t and back.k == k and back.n_iter == 5
            _assert_same(est_ref.get_log(), back.get_log(), 'reloaded get_log()')
            for attr in log_attrs + ['sum_f', 'sum_diff_f', 'vecsum_f_adaptive', 'sum_diff_f_exact']:
                _assert_same(getattr(est_ref, attr), getattr(back, attr), 'reloaded ' + attr)
            for log_name in est_ref.list_log_names:
                a, b = est_ref.stratif[log_name], back.stratif[log_name]
                if a.n_adaptive_strata:
                    np.testing.assert_allclose(b.adaptive_rel_alloc, a.adaptive_rel_alloc, rtol=1e-9, atol=1e-15)
                    np.testing.assert_allclose(b.guess_rel_alloc, a.guess_rel_alloc, rtol=1e-9, atol=1e-15)
        # a reloaded estimator goes on exactly where the saved one stopped
        back = aslsd_b200.MHPStratEstimB200(engine=OracleEngine())
        back.load(file_new, mhp, path)
        state = copy.deepcopy(rng_ref.bit_generator.state)
        g_ref = est_ref.lse_k_grad_estimate(x, rng=rng_ref, grad_alloc=False)
        rng_back = np.random.default_rng(0)
        rng_back.bit_generator.state = state
        g_back = back.lse_k_grad_estimate(x, rng=rng_back, grad_alloc=False)
        np.testing.assert_allclose(g_back, g_ref, rtol=1e-10, atol=1e-10 * np.abs(g_ref).max())
        assert rng_ref.bit_generator.state == rng_back.bit_generator.state


@pytest.mark.parametrize('name,T_f', [('exp1d', 300.), ('gauss1d', 100.)])
def test_exact_estimator_logs_match_reference(ref, name, T_f, tmp_path):
    """is_log_total_estimates and save / load of the exact estimator (mhp_exact_estimator.py:272-330,
    :793-935).  One-dimensional cases: upstream keeps only the last phi stratification
    (the assignment at :101 is outside its loop), so for d >= 2 its own log initialisation and save
    raise KeyError; the d = 2 round trip of the B200 estimator is checked below."""
    import aslsd_b200
    model, mhp, path = _simulate(ref, name, T_f)
    engine = OracleEngine()
    for k in range(model['d']):
        est_ref = ref.MHPExactEstim(is_log_lse=True, is_log_total_estimates=True)
        est_ref.initialize(k, 3, mhp, path)
        est_new = aslsd_b200.MHPExactEstimB200(is_log_lse=True, is_log_total_estimates=True, engine=engine)
        est_new.initialize(k, 3, mhp, path)
        x = np.array(case_x(name)[k], dtype=object)
        for t in range(3):
            est_ref.lse_k_grad_estimate(x)
            est_new.lse_k_grad_estimate(x)
            x = x * (1. + 0.01 * np.cos(t + np.arange(len(x))))
        _assert_same(est_ref.get_log(), est_new.get_log(), 'get_log()')
        log_attrs = [a for a in vars(est_ref) if a.endswith('_log')]
        assert sorted(log_attrs) == sorted(a for a in vars(est_new) if a.endswith('_log'))
        for attr in log_attrs:
            _assert_same(getattr(est_ref, attr), getattr(est_new, attr), attr)
        # save / load (mhp_exact_estimator.py:793-935; the stratification files the reference's
        # load expects are written too)
        file_new = str(tmp_path / ('new_%d.pickle' % k))
        est_new.save(file_new)
        for make in (lambda: aslsd_b200.MHPExactEstimB200(engine=engine), ref.MHPExactEstim):
            back = make()
            back.load(file_new, mhp, path)
            assert back.t == 3 and back.k == k
            _assert_same(est_ref.get_log(), back.get_log(), 'reloaded get_log()')
            for attr in log_attrs + ['sum_f', 'sum_diff_f']:
                _assert_same(getattr(est_ref, attr), getattr(back, attr), 'reloaded ' + attr)


def test_exact_estimator_save_load_two_dimensions(ref, tmp_path):
    import aslsd_b200
    model, mhp, path = _simulate(ref, 'exp2d', 60.)
    engine = OracleEngine()
    for k in range(2):
        est = aslsd_b200.MHPExactEstimB200(is_log_lse=True, is_log_total_estimates=True, engine=engine)
        est.initialize(k, 2, mhp, path)
        x = np.array(case_x('exp2d')[k], dtype=object)
        g = est.lse_k_grad_estimate(x)
        file = str(tmp_path / ('exact_%d' % k))      # no .pickle suffix: the other naming branch
        est.save(file)
        assert os.path.exists(file + '_strf_phi_0') and os.path.exists(file + '_current_state')
        back = aslsd_b200.MHPExactEstimB200(engine=engine)
        back.load(file, mhp, path)
        _assert_same(est.get_log(), back.get_log(), 'get_log()')
        _assert_same(est.sum_diff_f, back.sum_diff_f, 'sum_diff_f')
        _assert_same(est.sum_diff_f_log, back.sum_diff_f_log, 'sum_diff_f_log')
        np.testing.assert_allclose(back.recombine_grad_lse_k(), g, rtol=1e-13)
        np.testing.assert_allclose(back.lse_k_grad_estimate(x), g, rtol=1e-12)


def test_per_term_entry_points_match_reference(ref):
    """compute_{psi,upsilonzero,phi,sim_upsilon,cross_upsilon}_sum called one by one
    (mhp_exact_estimator.py:543-707, mhp_stratified_estimator.py:1081-1331): same sums in the same
    attributes and, for the stratified estimator, the same draws from the shared Generator."""
    import aslsd_b200
    name = 'exp2d'
    model, mhp, path = _simulate(ref, name, 150.)
    k = 1
    x = np.asarray(case_x(name)[k], dtype=float)
    x_k0, x_k1, mu_k = x[1:3], x[3:5], x[0]

    def run(est, **kw):
        est.clear_estimates()
        est.compute_psi_sum(0, x_k0, mu_k, **kw)
        est.compute_phi_sum(1, x_k1, **kw)
        est.compute_cross_upsilon_sum(0, 1, x_k0, x_k1, **kw)
        est.compute_sim_upsilon_sum(1, x_k1, **kw)
        est.compute_upsilonzero_sum(0, x_k0, **kw)

    touched = ['psi_0', 'phi_1', 'upsilon_0_1', 'upsilon_1_1', 'upsilonzero_0']
    # exact
    a = ref.MHPExactEstim()
    a.initialize(k, 2, mhp, path)
    b = aslsd_b200.MHPExactEstimB200(engine=OracleEngine())
    b.initialize(k, 2, mhp, path)
    run(a)
    run(b)
    for log_name in a.list_log_names:
        _assert_same(a.sum_f[log_name], b.sum_f[log_name], 'sum_f ' + log_name)
        _assert_same(a.sum_diff_f[log_name], b.sum_diff_f[log_name], 'sum_diff_f ' + log_name)
        if log_name not in touched:
            assert b.sum_f[log_name] == 0.
    # stratified
    a = ref.MHPStratEstim()
    a.initialize(k, 3, mhp, path, **STRAT_KW)
    b = aslsd_b200.MHPStratEstimB200(engine=OracleEngine())
    b.initialize(k, 3, mhp, path, **STRAT_KW)
    rng_a, rng_b = np.random.default_rng(21), np.random.default_rng(21)
    for _ in range(2):
        run(a, rng=rng_a, grad_alloc=False, compute_f_sum=True)
        run(b, rng=rng_b, grad_alloc=False, compute_f_sum=True)
        assert rng_a.bit_generator.state == rng_b.bit_generator.state
        for log_name in a.list_log_names:
            for attr in ('sum_f', 'sum_diff_f', 'vecsum_f_adaptive', 'sum_f_nonadaptive', 'sum_f_exact',
                         'sum_diff_f_adaptive', 'sum_diff_f_nonadaptive', 'sum_diff_f_exact'):
                _assert_same(getattr(a, attr)[log_name], getattr(b, attr)[log_name], attr + ' ' + log_name)
            if a.stratif[log_name].n_adaptive_strata:
                np.testing.assert_allclose(b.stratif[log_name].adaptive_rel_alloc,
                                           a.stratif[log_name].adaptive_rel_alloc, rtol=1e-9, atol=1e-15)


def test_reference_model_save_writes_the_b200_estimators(ref, tmp_path):
    """MHP.save (mhp.py:2380-2437) saves fit_estim[k] inside a bare try/except: an estimator whose
    save raises is silently dropped.  After a fit with the B200 estimators the per-k files are there
    and the reference's own estimator class reads them back."""
    import aslsd_b200
    model, mhp, path = _simulate(ref, 'exp2d', 120.)
    x_0 = [np.array(x, dtype=object) for x in case_x('exp2d')]
    mhp.fit(path, x_0=copy.deepcopy(x_0), n_iter=3, seed=99, write=True,
            estimators=aslsd_b200.MHPStratEstimB200(engine=OracleEngine(), is_log_lse=True), **STRAT_KW)
    assert isinstance(mhp.fit_estim[0], aslsd_b200.MHPStratEstimB200)
    prefix = str(tmp_path / 'model')
    mhp.save(prefix)
    for k in range(2):
        file_k = prefix + '_estimator_%d' % k
        assert os.path.exists(file_k + '_basic_attr') and os.path.exists(file_k + '_strf_upsilon_0_1')
        back = ref.MHPStratEstim()
        back.load(file_k, mhp, path)
        assert back.t == 3 and back.k == k
        np.testing.assert_array_equal(back.logged_lse, mhp.fit_estim[k].logged_lse)
        _assert_same(back.sum_diff_f, mhp.fit_estim[k].sum_diff_f, 'sum_diff_f')
