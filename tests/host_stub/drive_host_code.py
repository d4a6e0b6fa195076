"""Driver of tests/test_host_code_sanitizer.py: runs in a subprocess with the AddressSanitizer
runtime preloaded and the library built against the do-nothing CUDA runtime
(tests/host_stub/cudart_stub.cpp).  Every public engine call is made on several models; kernel
launches are no-ops, so outputs are meaningless - what is checked is that the host code (argument
checks, planning, packing, assembly, buffer management) runs clean and returns the documented
error codes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402

from aslsd_b200 import capi  # noqa: E402

capi.LIB_PATH = sys.argv[1]
capi._lib = None
import aslsd_b200  # noqa: E402
from aslsd_b200.distributed import block_slices, time_block_ranges  # noqa: E402
from cases import CASES, MTLH_CASES, case_x, mtlh_inputs  # noqa: E402


def exercise(name, engine):
    model = CASES[name][0]
    d = model['d']
    rng = np.random.default_rng(3)
    T_f = 300.
    times = [np.sort(rng.uniform(0, T_f, size=400 + 150 * i)) for i in range(d)]
    for i in range(d):      # a few targets close to T_f: boundary work items
        times[i][-3:] = T_f - np.array([1., .1, .001]) * (1 + i)
        times[i].sort()
    engine.set_events(times, T_f)
    engine.index_maps()
    engine.set_model(model)
    xs = case_x(name)
    ks = list(range(d))
    for k in ks:
        engine.exact_sums(k, xs[k])
        engine.exact_sums(k, xs[k], compute_f_sum=True, compute_diff_f_sum=False)
        engine.exact_sums_packed(k, xs[k], count_pairs=True)
        engine.point_sums(k, xs[k], 'phi')
        engine.point_sums(k, xs[k], 'psi')
    engine.exact_sums_batch(ks, xs)
    work = engine.last_work_counts()
    # the plan itself is host work: items of the d sub-problems before sharing, unique rows after
    R = len(model['kernels'][0][0])
    assert work['items'] == d * (d * (2 * R + 2 * R * R) + d * (d - 1) * R * R), (name, work)
    assert work['rows'] <= work['items']
    if name in ('sbf2d3', 'sbf3d2'):
        # (beta, delta) fixed: upsilon, upsilonzero and psi rows do not depend on k, phi rows do
        assert work['rows'] == d * d * R * R + d * R * R + d * R + d * d * R, (name, work)
    engine.exact_sums_batch(ks, xs, count_pairs=True)
    engine.exact_max_lag(ks, xs)
    for block in time_block_ranges(times, 3):
        engine.exact_sums_batch(ks, xs, ranges=block)
        local, ranges = block_slices(times, block, engine.exact_max_lag(ks, xs))
    try:
        engine.use_lag_quadrature(False)
        engine.exact_sums_batch(ks, xs)
    finally:
        engine.use_lag_quadrature(True)
    # raw rows on the "device", assembled from a host copy
    n_raw = engine.exact_raw_len_batch(d)
    raw = np.zeros(n_raw)
    engine.exact_raw_device_batch(ks, xs, raw.ctypes.data)
    engine.exact_assemble_batch(ks, xs, raw[:engine.exact_raw_used()])
    # sampled sums: every term kind, several strata
    for k in ks:
        terms, offsets, ms, ns = [], [0], [], []
        for kind, i, j in [(0, 0, 0), (1, d - 1, d - 1), (2, 0, 0), (2, d - 1, d - 1), (3, 0, 0), (3, 0, d - 1),
                           (3, d - 1, 0)]:
            tgt, src = (k, i) if kind == 2 else (i, j)
            terms.append((kind, i, j, 3))
            for s in range(3):
                n_s = int(rng.integers(1, 60))
                m = rng.integers(1, len(times[tgt]), size=n_s)
                hi = np.maximum(np.searchsorted(times[src], times[tgt][m], side='left'), 1)
                n = np.maximum(hi - rng.integers(1, 30, size=n_s), 0) if kind >= 2 else np.zeros(n_s, dtype=np.int64)
                ms.append(m)
                ns.append(n)
                offsets.append(offsets[-1] + n_s)
        n_strata = len(offsets) - 1
        for grad_alloc in (False, True):
            engine.sampled_sums(k, xs[k], terms, offsets, np.concatenate(ms), np.concatenate(ns),
                                strata_sizes=np.full(n_strata, 50.), default_alloc=np.full(n_strata, .1),
                                compute_diff_f_sum=True, grad_alloc=grad_alloc)
        # documented argument errors
        bad_m = np.concatenate(ms).copy()
        bad_m[0] = 10 ** 9
        for call in (lambda: engine.sampled_sums(k, xs[k], terms, offsets, bad_m, np.concatenate(ns)),
                     lambda: engine.exact_sums(k, xs[k][:-1]),
                     lambda: engine.exact_sums(d + 3, xs[k])):
            try:
                call()
            except (aslsd_b200.EngineError, ValueError, IndexError):
                pass
            else:
                raise AssertionError('bad arguments were accepted')


def exercise_estimators(name, engine):
    """The estimator classes end to end (their host side: sampling, packing, result attributes)."""
    model = CASES[name][0]
    d = model['d']
    rng = np.random.default_rng(5)
    T_f = 200.
    times = [np.sort(rng.uniform(0, T_f, size=300 + 100 * i)) for i in range(d)]
    # the stub computes no index maps: the samplers get the oracle's (test infrastructure)
    from oracle import time_ordering as OT
    kappa, varpi = OT.kappa_varpi_closed_form(times, T_f)
    path = aslsd_b200.ProcessPath(times, T_f, kappa=kappa, varpi=varpi, lag_sizes=OT.lag_sizes(times, kappa, varpi),
                                  engine=engine)
    mirror = aslsd_b200.MHP(model)
    kw = dict(n_exact_single=20, n_samples_adaptive_single=30, nonadaptive_sample_size_single=40,
              n_samples_adaptive_double=100, nonadaptive_sample_size_double=60)
    for k in range(d):
        x = case_x(name)[k]
        exact = aslsd_b200.MHPExactEstimB200(engine=engine, is_log_lse=True)
        exact.initialize(k, 3, mirror, path)
        exact.lse_k_grad_estimate(x)
        exact.lse_k_estimate(x)
        strat = aslsd_b200.MHPStratEstimB200(engine=engine, is_log_lse=True, is_log_allocs=True, is_log_ixs=True)
        strat.initialize(k, 3, mirror, path, **kw)
        for grad_alloc in (False, True):
            strat.lse_k_grad_estimate(x, rng=rng, grad_alloc=grad_alloc)
    engine.last_work_counts()
    engine.last_timing()
    engine.index_maps_kernel_time()
    engine.update_allocation(np.array([10., 20., 0.]), np.array([1., 2., 0.]), np.array([.3, .3, .4]))
    for option in (True, False):
        engine.use_weight_cache(option)
        engine.exact_sums_batch(list(range(d)), case_x(name))
    engine.use_weight_cache(True)


def exercise_marked(name, engine):
    model, xs, n_events, T_f = MTLH_CASES[name]
    times, marks, T_f = mtlh_inputs(name)
    mirror = aslsd_b200.MTLH(model)
    from oracle import time_ordering as OT
    kappa, varpi = OT.kappa_varpi_closed_form(times, T_f)
    path = aslsd_b200.ProcessPath(times, T_f, list_marks=marks, kappa=kappa, varpi=varpi,
                                  lag_sizes=OT.lag_sizes(times, kappa, varpi), engine=engine)
    for k in range(model['d']):
        x = np.asarray(xs[k], dtype=float)
        est = aslsd_b200.MTLHExactEstimB200(engine=engine)
        est.initialize(k, 2, mirror, path)
        est.compute_objective(x, compute_f_sum=True, compute_diff_f_sum=True)
        if model['impacts'][0][0]['type'] == 'linear' and 'baselines' not in model:
            strat = aslsd_b200.MTLHStratEstimB200(engine=engine)
            strat.initialize(k, 2, mirror, path, n_exact_single=10, n_samples_adaptive_single=20,
                             nonadaptive_sample_size_single=20, n_samples_adaptive_double=40,
                             nonadaptive_sample_size_double=30)
            strat.lse_k_grad_estimate(x, rng=np.random.default_rng(1), grad_alloc=False)


def main():
    engine = aslsd_b200.Engine(0)
    for name in ('exp1d', 'exp2d', 'gauss2d', 'sbf2d3', 'mix2d', 'exp1d2b', 'gauss1d2b', 'exp3d', 'sbf3d2',
                 'gauss3d'):
        exercise(name, engine)
    for name in ('exp2d', 'sbf2d3', 'mix1d', 'gauss3d'):
        exercise_estimators(name, engine)
    engine.close()
    engine = aslsd_b200.Engine(0)
    for name in ('mtlh_exp2d', 'mtlh_gauss1d', 'mtlh_exp1d_m2', 'mtlh_exp2d_const', 'mtlh_exp3d', 'mtlh_exp2d_pc',
                 'mtlh_gauss1d_pc'):
        exercise_marked(name, engine)
    # several contexts side by side; unsorted events and NaNs are rejected
    other = aslsd_b200.Engine(0)
    other.set_events([np.array([1., 2., 3.])], 5.)
    for bad in ([np.array([2., 1., 3.])], [np.array([1., np.nan, 3.])]):
        try:
            other.set_events(bad, 5.)
        except (aslsd_b200.EngineError, ValueError):
            pass
        else:
            raise AssertionError('invalid events were accepted')
    other.close()
    # the error-path tests of the device suite are host-side argument checks: they run here as they are
    import test_gpu_exact
    import test_gpu_stratified
    test_gpu_exact.test_errors(engine)
    test_gpu_stratified.test_sampled_sums_reject_bad_indices(engine)
    # simulation: spawn tables, pool sizing, the generation loop's bookkeeping and the error codes
    # (the stub's kernels draw nothing, so no event is ever born)
    import test_gpu_simulate
    test_gpu_simulate.test_simulation_rejects_unstable_models(engine)
    for name in ('exp2d', 'sbf2d3', 'mix2d'):
        model, mu, kp, _ = CASES[name]
        times = engine.simulate(model, mu, kp, 50., seed=5)
        assert len(times) == model['d']
    engine.close()
    print('host code ok')


if __name__ == '__main__':
    main()
