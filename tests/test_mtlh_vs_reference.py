"""MTLHExactEstimB200 against the live reference (CPU, oracle-backed engine double): the host
logic of the marked-model estimator - impact evaluation by the reference's own closures, weight
roles, term masks, nested result blocks, recombination - and the plug-in into MTLH.fit."""
import copy

import numpy as np
import pytest

import ref_shim
from cases import MTLH_CASES, mtlh_inputs
from oracle import mtlh_exact as OM
from oracle_engine import OracleEngine

pytestmark = pytest.mark.reference


def _reference_setup(name):
    import make_golden_mtlh
    ref_shim.make_importable()
    from aslsd.stats.events.process_path import ProcessPath
    model, xs, n_events, T_f = MTLH_CASES[name]
    mtlh = make_golden_mtlh.build_reference_mtlh(model)
    times, marks, T_f = mtlh_inputs(name)
    return model, xs, mtlh, ProcessPath(times, T_f, list_marks=marks)


@pytest.mark.parametrize('name', ['mtlh_exp2d', 'mtlh_gauss1d', 'mtlh_exp1d_m2', 'mtlh_exp3d', 'mtlh_exp2d_pc',
                                  'mtlh_gauss1d_pc'])
def test_estimator_state_matches_reference_estimator(name):
    import aslsd_b200
    from aslsd.optimize.estimators.mtlh_exact_estimator import MTLHExactEstim
    model, xs, mtlh, path = _reference_setup(name)
    for k in range(model['d']):
        x = np.asarray(xs[k], dtype=float)
        ref_est = MTLHExactEstim(is_log_lse=True)
        ref_est.initialize(k, 3, mtlh, path)
        new_est = aslsd_b200.MTLHExactEstimB200(is_log_lse=True, engine=OracleEngine())
        new_est.initialize(k, 3, mtlh, path)
        g_ref = ref_est.lse_k_grad_estimate(x)
        g_new = new_est.lse_k_grad_estimate(x)
        a = OM.flatten(ref_est.M_term, ref_est.diff_M_term, ref_est.sum_f, ref_est.sum_diff_f, 'k')
        b = OM.flatten(new_est.M_term, new_est.diff_M_term, new_est.sum_f, new_est.sum_diff_f, 'k')
        assert sorted(a) == sorted(b)
        for key in a:
            np.testing.assert_allclose(b[key], a[key], rtol=1e-10, atol=1e-12, err_msg=key)
        np.testing.assert_allclose(g_new, g_ref, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(new_est.logged_lse, ref_est.logged_lse, rtol=1e-10)
        np.testing.assert_allclose(new_est.lse_k_estimate(x), ref_est.lse_k_estimate(x), rtol=1e-10)


def test_plugs_into_reference_mtlh_fit():
    """A single instance passed to MTLH.fit(estimators=...) is deep-copied per dimension and
    drives the reference's own solver loop to the same iterates as its MTLHExactEstim."""
    import aslsd_b200
    from aslsd.optimize.estimators.estimator import Estimator
    assert issubclass(aslsd_b200.MTLHExactEstimB200, Estimator)
    model, xs, mtlh, path = _reference_setup('mtlh_exp2d')
    x_0 = [np.asarray(x, dtype=float) for x in xs]
    mu_ref, ker_ref, imp_ref = mtlh.fit(path, x_0=copy.deepcopy(x_0), n_iter=4, exact_grad=True, seed=3)
    import make_golden_mtlh
    mtlh2 = make_golden_mtlh.build_reference_mtlh(model)
    mu_new, ker_new, imp_new = mtlh2.fit(path, x_0=copy.deepcopy(x_0), n_iter=4, seed=3,
                                         estimators=aslsd_b200.MTLHExactEstimB200(engine=OracleEngine()))
    flat = lambda mu, ker, imp: np.concatenate(
        [np.ravel(np.asarray(v, dtype=float)) for v in list(mu) + [c for row in ker for c in row]
         + [c for row in imp for c in row]])
    np.testing.assert_allclose(flat(mu_new, ker_new, imp_new), flat(mu_ref, ker_ref, imp_ref), rtol=1e-8)


def test_stratified_estimator_plugs_into_reference_mtlh_fit():
    """MTLH.fit's default is the stratified estimator (mtlh.py:2379-2386): the B200 counterpart
    follows the same iterates (same samples from the shared Generator, same ADAM steps)."""
    import aslsd_b200
    import make_golden_mtlh
    import make_golden_mtlh_strat as MS
    from aslsd.stats.events.process_path import ProcessPath
    model, times, marks, T_f, kw, n_iter, seed = MS.strat_inputs('mtlh_exp2d')
    path = ProcessPath(times, T_f, list_marks=marks)
    xs = MTLH_CASES['mtlh_exp2d'][1]
    x_0 = [np.asarray(x, dtype=float) for x in xs]
    ref_model = make_golden_mtlh.build_reference_mtlh(model)
    mu_ref, ker_ref, imp_ref = ref_model.fit(path, x_0=copy.deepcopy(x_0), n_iter=5, seed=9, **kw)
    new_model = make_golden_mtlh.build_reference_mtlh(model)
    mu_new, ker_new, imp_new = new_model.fit(
        path, x_0=copy.deepcopy(x_0), n_iter=5, seed=9,
        estimators=aslsd_b200.MTLHStratEstimB200(engine=OracleEngine()), **kw)
    flat = lambda mu, ker, imp: np.concatenate(
        [np.ravel(np.asarray(v, dtype=float)) for v in list(mu) + [c for row in ker for c in row]
         + [c for row in imp for c in row]])
    np.testing.assert_allclose(flat(mu_new, ker_new, imp_new), flat(mu_ref, ker_ref, imp_ref), rtol=1e-8)


def test_unsupported_baselines_raise():
    """Baselines other than ConstantBaseline / PieceConstBaseline are outside the device's scope (the
    reference's exact estimator cannot evaluate them either): loud refusal; the stratified estimator
    refuses PieceConstBaseline as well."""
    import types
    import aslsd_b200
    model, xs, mtlh, path = _reference_setup('mtlh_exp1d_m2')
    other = type('CosineBaseline', (), {})()
    mtlh._baselines_vec = [types.SimpleNamespace(_basis_mus=[other])]
    est = aslsd_b200.MTLHExactEstimB200(engine=OracleEngine())
    with pytest.raises(NotImplementedError):
        est.initialize(0, 3, mtlh, path)
    model, xs, mtlh, path = _reference_setup('mtlh_gauss1d_pc')
    with pytest.raises(NotImplementedError):
        aslsd_b200.MTLHStratEstimB200(engine=OracleEngine()).initialize(0, 3, mtlh, path)


def test_marked_estimators_save_load(tmp_path):
    """save / load in the reference's file layout (mtlh_exact_estimator.py:1058-1209,
    mtlh_stratified_estimator.py:1749-1929).  Upstream's own save raises AttributeError for marked
    models (it lists mu_term, which those estimators never set), so this is a round trip of the
    B200 estimators: the reloaded estimator holds the same state and goes on with the same draws."""
    import aslsd_b200
    model, xs, mtlh, path = _reference_setup('mtlh_exp2d')
    x = np.asarray(xs[0], dtype=float)
    for cls, kw in ((aslsd_b200.MTLHExactEstimB200, dict(is_log_lse=True, is_log_total_estimates=True)),
                    (aslsd_b200.MTLHStratEstimB200, dict(is_log_lse=True, is_log_allocs=True, is_log_ixs=True))):
        est = cls(engine=OracleEngine(), **kw)
        est.initialize(0, 4, mtlh, path)
        rng = np.random.default_rng(4)
        g = est.lse_k_grad_estimate(x, rng=rng, grad_alloc=False)
        file = str(tmp_path / (cls.__name__ + '.pickle'))
        est.save(file)
        back = cls(engine=OracleEngine())
        back.load(file, mtlh, path)
        assert back.t == 1 and back.k == 0 and back.n_iter == 4
        np.testing.assert_array_equal(back.logged_lse, est.logged_lse)
        a = OM.flatten(est.M_term, est.diff_M_term, est.sum_f, est.sum_diff_f, 'k')
        b = OM.flatten(back.M_term, back.diff_M_term, back.sum_f, back.sum_diff_f, 'k')
        assert sorted(a) == sorted(b)
        for key in a:
            np.testing.assert_array_equal(b[key], a[key], err_msg=key)
        np.testing.assert_allclose(back.recombine_grad_lse_k(), g, rtol=1e-13)
        rng_back = np.random.default_rng(0)
        rng_back.bit_generator.state = copy.deepcopy(rng.bit_generator.state)
        g_next = est.lse_k_grad_estimate(x, rng=rng, grad_alloc=False)
        g_back = back.lse_k_grad_estimate(x, rng=rng_back, grad_alloc=False)
        np.testing.assert_allclose(g_back, g_next, rtol=1e-12)
        assert rng.bit_generator.state == rng_back.bit_generator.state
