"""Marked models (MTLH) on the device: MTLHExactEstimB200 against the golden vectors generated
from the reference's MTLHExactEstim (every per-term sum and derivative block, M term, LSE,
gradient), against the oracle at a larger size, and the weighted C-ABI entry point itself.

Tolerance: relative 1e-10 per entry (BASELINE.json north_star), fp64."""
import numpy as np
import pytest

from cases import MTLH_CASES, MTLH_NAMES_2D, mtlh_inputs
from helpers import RTOL, load_golden
from oracle import mtlh_exact as OM

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
    import aslsd_b200
    return aslsd_b200.Engine(0)


def _flat_state(est, prefix):
    return OM.flatten(est.M_term, est.diff_M_term, est.sum_f, est.sum_diff_f, prefix)


@pytest.mark.parametrize('name', MTLH_NAMES_2D)
def test_mtlh_exact_matches_reference_golden(engine, name):
    import aslsd_b200
    model, xs, n_events, T_f = MTLH_CASES[name]
    g = load_golden(name)
    times, marks, T_f = mtlh_inputs(name)
    mirror = aslsd_b200.MTLH(model)
    path = aslsd_b200.ProcessPath(times, T_f, list_marks=marks, engine=engine)
    for k in range(model['d']):
        est = aslsd_b200.MTLHExactEstimB200(is_log_lse=True, engine=engine)
        est.initialize(k, 5, mirror, path)
        x = g['k%d/x' % k]
        lse = est.lse_k_estimate(x, compute_diff_f_sum=True, count_iter=False)
        flat = _flat_state(est, 'k%d' % k)
        for key, val in flat.items():
            ref = g[key]
            np.testing.assert_allclose(val, ref, rtol=RTOL, atol=1e-12 * np.max(np.abs(ref), initial=1.),
                                       err_msg='%s %s' % (name, key))
        np.testing.assert_allclose(lse, g['k%d/lse' % k], rtol=RTOL)
        grad = est.lse_k_grad_estimate(x)
        if 'k%d/grad' % k in g:
            ref = g['k%d/grad' % k]
            np.testing.assert_allclose(grad, ref, rtol=RTOL, atol=1e-10 * np.max(np.abs(ref)))
        assert est.t == 1 and abs(est.logged_lse[1] - g['k%d/lse' % k]) <= RTOL * abs(g['k%d/lse' % k])


def test_mtlh_exact_matches_oracle_medium(engine):
    """A few thousand events per type, 2 mark dimensions: against the literal oracle."""
    import aslsd_b200
    from cases import exp_kernel
    d = 2
    model = {'d': d, 'kernels': [[exp_kernel() for _ in range(d)] for _ in range(d)],
             'impacts': [[{'type': 'linear', 'mark_dim': 2} for _ in range(d)] for _ in range(d)]}
    rng = np.random.default_rng(5)
    T_f = 900.
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (1500, 1100)]
    marks = [rng.uniform(0, 1, size=(len(t), 2)) for t in times]
    mirror = aslsd_b200.MTLH(model)
    path = aslsd_b200.ProcessPath(times, T_f, list_marks=marks, engine=engine)
    x = np.array([.7, .2, 1.1, .3, 1.7, .4, .2, .1, .6])
    for k in range(d):
        est = aslsd_b200.MTLHExactEstimB200(engine=engine)
        est.initialize(k, 5, mirror, path)
        est.lse_k_estimate(x, compute_diff_f_sum=True)
        M, dM, sum_f, sum_diff_f = OM.exact_sums_literal(model, k, x, times, marks, T_f)
        got, want = _flat_state(est, 'k'), OM.flatten(M, dM, sum_f, sum_diff_f, 'k')
        for key in want:
            np.testing.assert_allclose(got[key], want[key], rtol=RTOL,
                                       atol=1e-12 * np.max(np.abs(want[key]), initial=1.), err_msg=key)
        np.testing.assert_allclose(est.recombine_grad_lse_k(),
                                   OM.recombine_grad_lse_k(model, k, dM, sum_diff_f, T_f),
                                   rtol=RTOL, atol=1e-12)


def test_weighted_lag_quadrature_agrees_with_item_kernels(engine):
    """Marked models go through the lag quadrature too (weighted moments): same sums as the
    per-item pair kernels, and the moment pass really ran."""
    import aslsd_b200
    from cases import sbf_gauss_kernel
    d = 2
    model = {'d': d, 'kernels': [[sbf_gauss_kernel(3) for _ in range(d)] for _ in range(d)],
             'impacts': [[{'type': 'linear', 'mark_dim': 1} for _ in range(d)] for _ in range(d)]}
    rng = np.random.default_rng(9)
    T_f = 5000.
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (6000, 4500)]
    marks = [rng.uniform(0, 1, size=(len(t), 1)) for t in times]
    mirror = aslsd_b200.MTLH(model)
    path = aslsd_b200.ProcessPath(times, T_f, list_marks=marks, engine=engine)
    x = np.array([.5, .2, .1, .1, .1, .2, .05, .6, .3])
    states = []
    try:
        for quadrature in (True, False):
            engine.use_lag_quadrature(quadrature)
            est = aslsd_b200.MTLHExactEstimB200(engine=engine)
            est.initialize(0, 3, mirror, path)
            est.lse_k_estimate(x, compute_diff_f_sum=True)
            work = engine.last_work_counts()
            assert (work['moment_pairs'] > 0) == quadrature
            states.append(_flat_state(est, 'k'))
    finally:
        engine.use_lag_quadrature(True)
    for key in states[0]:
        ref = states[1][key]
        np.testing.assert_allclose(states[0][key], ref, rtol=1e-11,
                                   atol=1e-12 * np.max(np.abs(ref), initial=1.), err_msg=key)


def test_unit_weights_reproduce_the_unmarked_sums(engine):
    """aslsd_exact_sums_weighted with all-ones weights equals aslsd_exact_sums (per-item kernels)."""
    from cases import CASES, case_x
    name = 'gauss2d'
    model = CASES[name][0]
    rng = np.random.default_rng(3)
    times = [np.sort(rng.uniform(0, 500., size=800)) for _ in range(2)]
    engine.set_events(times, 500.)
    engine.set_model(model)
    x = case_x(name)[0]
    engine.use_lag_quadrature(False)
    try:
        plain = engine.exact_sums_packed(0, x)
    finally:
        engine.use_lag_quadrature(True)
    engine.set_weights([(0, np.ones(800)), (1, np.ones(800))])
    roles = np.array([[0, 1]] * 5, dtype=np.int32)
    weighted = engine.exact_sums_weighted(0, x, roles)
    # (the weighted and the plain moment kernels associate their sums differently)
    np.testing.assert_allclose(weighted, plain, rtol=1e-12, atol=1e-300)
    # masked terms stay at zero, the others are unchanged
    from aslsd_b200.engine import term_layout
    layout, _ = term_layout(engine.desc, 0)
    mask = np.array([1 if entry[0].startswith('upsilon_0') else 0 for entry in layout], dtype=np.uint8)
    part = engine.exact_sums_weighted(0, x, roles, term_mask=mask)
    for (name_t, off, n1, n2, _), on in zip(layout, mask):
        seg = slice(off, off + 1 + n1 + n2)
        if on:
            np.testing.assert_allclose(part[seg], plain[seg], rtol=1e-13)
        else:
            assert np.all(part[seg] == 0.)
    with pytest.raises(Exception):
        engine.exact_sums_weighted(0, x, np.array([[1, 0]] * 5, dtype=np.int32))   # wrong event types


@pytest.mark.parametrize('name', ['mtlh_exp2d', 'mtlh_gauss1d'])
def test_mtlh_stratified_replays_reference_trace(engine, name):
    """MTLHStratEstimB200 on the device against traces of the reference's MTLHStratEstim: the
    samples are the reference's (same Generator stream to the last draw), every estimate within
    1e-10, the adaptive allocations within 1e-9."""
    from mtlh_replay import replay
    worst = replay(name, engine)
    assert worst <= RTOL
