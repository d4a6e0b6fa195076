"""Generate the golden vectors of tests/golden/ from the live reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case of tests/cases.py the reference simulates a path
(``MHP.simulate(seed=1234)``), builds its ``ProcessPath`` (kappa / varpi /
lag_sizes from ``time_ordering.get_kappa_varpi``) and evaluates
``MHPExactEstim.compute_objective(f + derivatives)``; the per-term sums, the
LSE value and the gradient (assembled with ``MHPStratEstim.recombine_grad_lse_k``,
SURVEY quirk 1) are stored next to the inputs.  ``index_maps.npz`` adds
tie-heavy paths and the reference's own 2x10 fixture
(stats/events/tests/time_ordering_test.py:70-78).  ``strat_*.npz`` hold
multi-iteration traces of ``MHPStratEstim`` (gradients, per-term sums and
adaptive allocations) for a fixed RNG seed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402
from cases import CASES, CASES_3D, case_x, flatten_sums  # noqa: E402


def kernel_param_array(kp, d):
    arr = np.empty((d, d), dtype=object)
    for a in range(d):
        for b in range(d):
            arr[a][b] = np.asarray(kp[a][b], dtype=float)
    return arr


def simulate(ref, mhp, model, mu, kp, T_f, seed=1234):
    d = model['d']
    return mhp.simulate(T_f, mu=np.asarray(mu, float),
                        kernel_param=kernel_param_array(kp, d), seed=seed,
                        verbose=False)


def make_exact(ref, names=None):
    for name, (model, mu, kp, T_f) in CASES.items():
        if names is not None and name not in names:
            continue
        d = model['d']
        mhp = ref_shim.build_reference_model(ref, model)
        path = simulate(ref, mhp, model, mu, kp, T_f)
        out = {'T_f': np.float64(T_f), 'd': np.int64(d)}
        for i in range(d):
            out['times/%d' % i] = np.asarray(path.list_times[i], dtype=float)
        for i in range(d):
            for j in range(d):
                out['kappa/%d/%d' % (i, j)] = np.asarray(path.kappa[i][j], np.int64)
                out['varpi/%d/%d' % (i, j)] = np.asarray(path.varpi[i][j], np.int64)
                out['lag_sizes/%d/%d' % (i, j)] = np.asarray(
                    [int(v) for v in path.lag_sizes[i][j]], np.int64)
        xs = case_x(name)
        for k in range(d):
            est = ref.MHPExactEstim()
            est.initialize(k, 10, mhp, path)
            est.compute_objective(np.array(xs[k], dtype=object),
                                  compute_f_sum=True, compute_diff_f_sum=True)
            out['k%d/x' % k] = xs[k]
            out.update(flatten_sums(est.sum_f, est.sum_diff_f, 'k%d' % k))
            out['k%d/lse' % k] = np.float64(est.recombine_lse_k())
            # the stratified estimator's assembly loop, applied to the exact
            # sums (the per-(p, r) body has the same text in both classes)
            est.recombine_diff_lse_theta_kpr = \
                est.recombine_diff_lse_theta_kpr_kernel
            out['k%d/grad' % k] = np.asarray(
                ref.MHPStratEstim.recombine_grad_lse_k(est), dtype=float)
            out['k%d/grad_exact_estim' % k] = np.asarray(
                est.recombine_grad_lse_k(), dtype=float)
        np.savez_compressed(os.path.join(HERE, 'exact_%s.npz' % name), **out)
        print('exact', name, [len(t) for t in path.list_times])


def make_index_maps(ref):
    out = {}
    rng = np.random.default_rng(2024)
    paths = {}
    # the reference's own fixture
    paths['ref_fixture'] = ([np.array([1.75869445, 1.80061029, 1.82564278,
                                       1.91613725, 1.92903655, 1.9636614,
                                       2.00255562, 2.0473686, 2.13208584,
                                       2.15969973]),
                             np.array([1.72669995, 1.75597784, 1.7805469,
                                       1.78599852, 1.79899688, 1.79990949,
                                       1.8040088, 2.02449236, 2.04648297,
                                       2.18191544])], 6.)
    # tie-heavy: times on a coarse grid, shared across dimensions
    for c in range(6):
        d = int(rng.integers(2, 5))
        times = []
        for i in range(d):
            n = int(rng.integers(3, 60))
            times.append(np.sort(rng.choice(np.arange(1, 80), size=n,
                                            replace=False)).astype(float) * 0.25)
        T_f = 25. if c % 2 == 0 else float(max(t[-1] for t in times))
        paths['ties_%d' % c] = (times, T_f)
    # unbalanced dimensions, continuous times
    for c in range(3):
        d = 3
        times = [np.sort(rng.uniform(0, 50, size=int(n)))
                 for n in rng.integers(1, 400, size=d)]
        paths['ragged_%d' % c] = (times, 50.)
    names = []
    for name, (times, T_f) in paths.items():
        kappa, varpi = ref.time_ordering.get_kappa_varpi(times, T_f)
        lags = ref.time_ordering.get_lag_sizes(times, kappa, varpi)
        d = len(times)
        names.append(name)
        out['%s/T_f' % name] = np.float64(T_f)
        out['%s/d' % name] = np.int64(d)
        for i in range(d):
            out['%s/times/%d' % (name, i)] = times[i]
            for j in range(d):
                out['%s/kappa/%d/%d' % (name, i, j)] = np.asarray(kappa[i][j], np.int64)
                out['%s/varpi/%d/%d' % (name, i, j)] = np.asarray(varpi[i][j], np.int64)
                out['%s/lag_sizes/%d/%d' % (name, i, j)] = np.asarray(
                    [int(v) for v in lags[i][j]], np.int64)
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'index_maps.npz'), **out)
    print('index maps', names)


if __name__ == '__main__':
    ref = ref_shim.import_reference()
    if sys.argv[1:] == ['3d']:     # only the three-type cases (added after the first set)
        make_exact(ref, names=set(CASES_3D))
        sys.exit(0)
    make_exact(ref)
    make_index_maps(ref)
    import make_golden_strat
    make_golden_strat.make_strat(ref)
    # the other generators (marked models, residuals, recursion values) are run the same way:
    #   python make_golden_mtlh.py; python make_golden_mtlh_strat.py;
    #   python make_golden_residuals.py; python make_golden_recexp.py
