"""Replay of the reference's MTLHStratEstim traces (tests/golden/strat_mtlh_*.npz) through
MTLHStratEstimB200 on a given engine: same evaluation points, one shared seeded Generator."""
import numpy as np

from cases import MTLH_CASES, MTLH_STRAT
from helpers import load_golden
from oracle import mtlh_exact as OM


def replay(name, engine, rtol=1e-10):
    import aslsd_b200
    g = load_golden('strat_' + name)
    case, n_events, T_f, kw, n_iter, seed = MTLH_STRAT[name]
    model = MTLH_CASES[case][0]
    d = model['d']
    times = [np.asarray(g['times/%d' % i]) for i in range(d)]
    marks = [np.asarray(g['marks/%d' % i]) for i in range(d)]
    mirror = aslsd_b200.MTLH(model)
    if hasattr(engine, 'index_maps'):
        path = aslsd_b200.ProcessPath(times, T_f, list_marks=marks, engine=engine)
    else:
        from helpers import _oracle_path
        path = _oracle_path(times, T_f)
        path.list_marks = marks
    rng = np.random.default_rng(seed)
    worst = 0.
    for k in range(d):
        est = aslsd_b200.MTLHStratEstimB200(is_log_lse=True, engine=engine)
        est.initialize(k, n_iter, mirror, path, **kw)
        names = [nm for nm in sorted(est.stratif) if est.stratif[nm].n_adaptive_strata > 0]
        for t in range(n_iter):
            x = g['k%d/t%d/x' % (k, t)]
            grad = est.lse_k_grad_estimate(x, rng=rng, grad_alloc=False)
            flat = OM.flatten(est.M_term, est.diff_M_term, est.sum_f, est.sum_diff_f, 'k%d/t%d' % (k, t))
            for key, val in flat.items():
                ref = g[key]
                scale = np.max(np.abs(ref), initial=0.)
                np.testing.assert_allclose(val, ref, rtol=rtol, atol=1e-12 * max(scale, 1.),
                                           err_msg='%s %s' % (name, key))
                worst = max(worst, float(np.max(np.abs(np.asarray(val) - ref)) / max(scale, 1e-300)))
            ref = g['k%d/t%d/grad' % (k, t)]
            np.testing.assert_allclose(grad, ref, rtol=rtol, atol=1e-10 * np.max(np.abs(ref)))
            np.testing.assert_allclose(est.lse_k, g['k%d/t%d/lse' % (k, t)], rtol=rtol)
            allocs = np.concatenate([np.asarray(est.stratif[nm].adaptive_rel_alloc, dtype=float)
                                     for nm in names])
            np.testing.assert_allclose(allocs, g['k%d/t%d/allocs' % (k, t)], rtol=1e-9, atol=1e-15,
                                       err_msg='%s k%d t%d allocations' % (name, k, t))
    assert str(rng.bit_generator.state) == str(g['rng_final']), 'RNG stream diverged from the reference'
    return worst
