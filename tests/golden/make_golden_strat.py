"""Golden traces of the reference's stratified (ASLSD) estimator.

Run in the build container only (needs /root/reference); called from
make_golden.py.  For every case the reference simulates a path, then for each
dimension k runs ``n_iter`` iterations of ``MHPStratEstim.lse_k_grad_estimate``
+ the reference's ADAM step, all k sharing one numpy Generator (as in MHP.fit,
models/hawkes/linear/mhp.py:1683-1787).  Stored per iteration: the evaluation
point, the gradient, every per-term sum, the LSE estimate and every adaptive
allocation.  A consumer that feeds the stored evaluation points to its own
estimator with the same seed must reproduce them; the integer sample
allocations are implied (they drive the RNG stream).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (HERE, os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE))):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from cases import CASES, case_x, lookup_case  # noqa: E402,F401

STRAT_CASES = {
    # name: (model case, T_f, n_iter, grad_alloc, stratification keywords)
    'strat_exp1d': ('exp1d', 2500., 40, False, dict()),     # default geometry of the reference
    'strat_exp2d': ('exp2d', 400., 25, False,
                    dict(n_exact_single=100, n_samples_adaptive_single=200,
                         nonadaptive_sample_size_single=300, n_samples_adaptive_double=600,
                         nonadaptive_sample_size_double=270)),
    'strat_exp2d_gradalloc': ('exp2d', 400., 25, True,
                              dict(n_exact_single=100, n_samples_adaptive_single=200,
                                   nonadaptive_sample_size_single=300, n_samples_adaptive_double=600,
                                   nonadaptive_sample_size_double=270)),
    'strat_gauss1d': ('gauss1d', 600., 25, True,
                      dict(n_exact_single=100, n_samples_adaptive_single=200,
                           nonadaptive_sample_size_single=300, n_samples_adaptive_double=600,
                           nonadaptive_sample_size_double=270)),
    'strat_sbf2d3': ('sbf2d3', 500., 20, False,
                     dict(n_exact_single=100, n_samples_adaptive_single=200,
                          nonadaptive_sample_size_single=300, n_samples_adaptive_double=600,
                          nonadaptive_sample_size_double=270)),
    # exponential + Gaussian basis in one kernel: the mixed upsilon closed forms, sampled
    'strat_mix1d': ('mix1d', 300., 20, False,
                    dict(n_exact_single=100, n_samples_adaptive_single=200,
                         nonadaptive_sample_size_single=300, n_samples_adaptive_double=600,
                         nonadaptive_sample_size_double=270)),
    # >= 100 iterations under the fit default grad_alloc=False ("identical integer
    # allocations over >= 100 iterations"): Gaussian kernels, where most single-index strata are
    # saturated and their allocation statistic is a rounding residue (aslsd_b200/saturation.py)
    'strat_gauss1d_long': ('gauss1d', 600., 120, False,
                           dict(n_exact_single=100, n_samples_adaptive_single=200,
                                nonadaptive_sample_size_single=300, n_samples_adaptive_double=600,
                                nonadaptive_sample_size_double=270)),
    'strat_sbf2d3_long': ('sbf2d3', 500., 100, False,
                          dict(n_exact_single=100, n_samples_adaptive_single=200,
                               nonadaptive_sample_size_single=300, n_samples_adaptive_double=600,
                               nonadaptive_sample_size_double=270)),
    # the benchmark's own model (ten event types, SBF-Gaussian(10)): every term family at d = 10, where most
    # cross terms upsilon_i_j have k != i, j; replayed on the CPU only (tests/test_stratified_golden.py)
    'strat_c5model': ('c5model', 250., 3, False,
                      dict(n_exact_single=40, n_samples_adaptive_single=80,
                           nonadaptive_sample_size_single=120, n_samples_adaptive_double=200,
                           nonadaptive_sample_size_double=90)),
    'strat_exp2d_long': ('exp2d', 400., 100, False,
                         dict(n_exact_single=100, n_samples_adaptive_single=200,
                              nonadaptive_sample_size_single=300, n_samples_adaptive_double=600,
                              nonadaptive_sample_size_double=270)),
}
SEED = 4321


def flat_terms(est):
    """Sorted-by-name flattening of sum_f / sum_diff_f."""
    f = np.array([est.sum_f[name] for name in sorted(est.sum_f)], dtype=float)
    parts = []
    for name in sorted(est.sum_diff_f):
        v = est.sum_diff_f[name]
        parts += [np.ravel(v[1]), np.ravel(v[2])] if isinstance(v, list) else [np.ravel(v)]
    return f, np.concatenate(parts).astype(float)


def flat_allocs(est):
    out = []
    for name in sorted(est.stratif):
        strf = est.stratif[name]
        if strf.n_adaptive_strata > 0:
            out.append(np.asarray(strf.adaptive_rel_alloc, dtype=float))
    return np.concatenate(out)


def make_strat(ref, only=None):
    from aslsd.optimize.solvers.adam import ADAM
    for name, (case, T_f, n_iter, grad_alloc, kw) in STRAT_CASES.items():
        if only and name not in only:
            continue
        model, mu, kp, _ = lookup_case(case)
        d = model['d']
        import ref_shim
        mhp = ref_shim.build_reference_model(ref, model)
        arr = np.empty((d, d), dtype=object)
        for a in range(d):
            for b in range(d):
                arr[a][b] = np.asarray(kp[a][b], dtype=float)
        path = mhp.simulate(T_f, mu=np.asarray(mu, float), kernel_param=arr, seed=1234,
                            verbose=False)
        out = {'T_f': np.float64(T_f), 'd': np.int64(d), 'n_iter': np.int64(n_iter),
               'grad_alloc': np.bool_(grad_alloc), 'seed': np.int64(SEED), 'case': np.array(case),
               'kw_names': np.array(sorted(kw)),
               'kw_values': np.array([kw[q] for q in sorted(kw)], dtype=np.int64)}
        for i in range(d):
            out['times/%d' % i] = np.asarray(path.list_times[i], dtype=float)
        rng = np.random.default_rng(SEED)
        xs = case_x(case)
        lower = [np.concatenate([[1e-10]] + [np.asarray(mhp.param_lower_bounds[k][i], float)
                                             for i in range(d)]) for k in range(d)]
        for k in range(d):
            est = ref.MHPStratEstim(is_log_lse=True)
            est.initialize(k, n_iter, mhp, path, **kw)
            solver = ADAM()
            x_k = np.array(xs[k], dtype=object)
            X = np.zeros((n_iter + 1, len(x_k)))
            G = np.zeros((n_iter, len(x_k)))
            F, DF, AL = [], [], []
            X[0] = x_k.astype(float)
            for t in range(n_iter):
                g = est.lse_k_grad_estimate(x_k, rng=rng, grad_alloc=grad_alloc)
                G[t] = g
                f, df = flat_terms(est)
                F.append(f)
                DF.append(df)
                AL.append(flat_allocs(est))
                x_k = np.clip(solver.iterate(t, x_k, g), lower[k], np.inf)
                X[t + 1] = x_k.astype(float)
            out['k%d/x' % k] = X
            out['k%d/grad' % k] = G
            out['k%d/sum_f' % k] = np.array(F)
            out['k%d/sum_diff_f' % k] = np.array(DF)
            out['k%d/allocs' % k] = np.array(AL)
            out['k%d/logged_lse' % k] = np.asarray(est.logged_lse, dtype=float)
        out['rng_final'] = np.array(str(rng.bit_generator.state))
        np.savez_compressed(os.path.join(HERE, '%s.npz' % name), **out)
        print('strat', name, [len(t) for t in path.list_times])


if __name__ == '__main__':
    import ref_shim
    make_strat(ref_shim.import_reference(), only=sys.argv[1:] or None)
