"""Golden traces of the reference's stratified estimator for marked models (MTLHStratEstim).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_mtlh_strat.py

For each case the reference model of ``cases.MTLH_CASES`` is evaluated on a seeded marked path for
a few iterations with one shared numpy Generator (``grad_alloc=False``, the only mode that runs
upstream): evaluation points, gradients, every per-term sum and derivative block, LSE, every
adaptive allocation after each call, and the final generator state are stored.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402
from cases import MTLH_CASES, MTLH_STRAT  # noqa: E402
from make_golden_mtlh import build_reference_mtlh  # noqa: E402
from oracle import mtlh_exact as OM  # noqa: E402


def strat_inputs(name):
    case, n_events, T_f, kw, n_iter, seed = MTLH_STRAT[name]
    model = MTLH_CASES[case][0]
    rng = np.random.default_rng(77)
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in n_events]
    q = max(imp.get('mark_dim', 1) for row in model['impacts'] for imp in row)
    marks = [rng.uniform(0, 1, size=(n, q)) for n in n_events]
    return model, times, marks, T_f, kw, n_iter, seed


def eval_points(name, k, n_iter):
    case = MTLH_STRAT[name][0]
    x = np.asarray(MTLH_CASES[case][1][k], dtype=float)
    rng = np.random.default_rng(100 + k)
    return [x * (1. + 0.05 * rng.uniform(-1, 1, size=len(x))) for _ in range(n_iter)]


def main():
    ref_shim.make_importable()
    warnings.simplefilter('ignore')
    from aslsd.stats.events.process_path import ProcessPath
    from aslsd.optimize.estimators.mtlh_stratified_estimator import MTLHStratEstim
    for name in MTLH_STRAT:
        model, times, marks, T_f, kw, n_iter, seed = strat_inputs(name)
        d = model['d']
        mtlh = build_reference_mtlh(model)
        path = ProcessPath(times, T_f, list_marks=marks)
        out = {'T_f': np.float64(T_f), 'd': np.int64(d), 'n_iter': np.int64(n_iter), 'seed': np.int64(seed)}
        for i in range(d):
            out['times/%d' % i] = times[i]
            out['marks/%d' % i] = marks[i]
        rng = np.random.default_rng(seed)
        for k in range(d):
            est = MTLHStratEstim(is_log_lse=True)
            est.initialize(k, n_iter, mtlh, path, **kw)
            names = [nm for nm in sorted(est.stratif) if est.stratif[nm].n_adaptive_strata > 0]
            for t, x in enumerate(eval_points(name, k, n_iter)):
                grad = est.lse_k_grad_estimate(x, rng=rng, grad_alloc=False)
                out['k%d/t%d/x' % (k, t)] = x
                out['k%d/t%d/grad' % (k, t)] = np.asarray(grad, dtype=float)
                out.update(OM.flatten(est.M_term, est.diff_M_term, est.sum_f, est.sum_diff_f,
                                      'k%d/t%d' % (k, t)))
                out['k%d/t%d/lse' % (k, t)] = np.float64(est.lse_k)
                out['k%d/t%d/allocs' % (k, t)] = np.concatenate(
                    [np.asarray(est.stratif[nm].adaptive_rel_alloc, dtype=float) for nm in names])
        out['rng_final'] = np.array(str(rng.bit_generator.state))
        np.savez_compressed(os.path.join(HERE, 'strat_%s.npz' % name), **out)
        print(name, [len(t) for t in times], 'lse k0', float(out['k0/t0/lse']))


if __name__ == '__main__':
    main()
