"""Timings of the LITERAL reference (saadlabyad/aslsd, imported from /root/reference through
tests/golden/ref_shim.py) on the host cores of the build container - BASELINE.md section 3, rows B1,
B3, B4.  The reference is pure Python and cannot travel to the GPU box, so these numbers are taken
here (core count recorded) and stored under profiles/; bench.py's reference arm times the numpy
oracle port instead.  One JSON line per measurement; single thread (the reference has no
parallelism).

    python scripts/time_literal_reference.py [B1] [B3] [B4] [B4x] > profiles/r1_literal_reference_cpu.jsonl
"""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np  # noqa: E402

import ref_shim  # noqa: E402

STRAT_KW = {}   # the fit defaults of the reference (mhp.py:1743-1750)


def emit(**kw):
    kw['host_cores'] = os.cpu_count()
    kw['threads'] = 1
    print(json.dumps(kw), flush=True)


def kernel_param_array(kp, d):
    arr = np.empty((d, d), dtype=object)
    for a in range(d):
        for b in range(d):
            arr[a][b] = np.asarray(kp[a][b], dtype=float)
    return arr


def configs():
    from aslsd_b200 import simulate
    exp = lambda: [{'type': 'exp'}]
    return {
        'C1': ({'d': 1, 'kernels': [[exp()]]}, [1.5], [[[.3, .8]]], 4.7e3),
        'C2': ({'d': 2, 'kernels': [[exp(), exp()], [exp(), exp()]]}, [1.5, 1.],
               [[[.2, 1.], [.6, 1.5]], [[.7, 2.], [.1, 1.3]]], 7.7e3),
        'C3': ({'d': 1, 'kernels': [[[{'type': 'gauss'}]]]}, [1.5], [[[.5, .5, 3.]]], 3.33e5),
        'C4': (simulate.sbf_gauss_model(1, 10), [1.0], [[simulate.random_mixture_weights(10, coeff=0.5)]], 5e5),
    }


def x_at(mu, kp, d):
    rng = np.random.default_rng(12)
    xs = []
    for k in range(d):
        x = np.concatenate([[mu[k]]] + [np.asarray(kp[k][i], float) for i in range(d)])
        xs.append(np.array(x * (1 + 0.1 * rng.uniform(-1, 1, size=len(x))), dtype=object))
    return xs


def reference_path(ref, model, mu, kp, T_f):
    """Events from the repo's simulator (same streams as the device benchmarks), book-keeping by
    the reference's own ProcessPath (Python bisect loops): timed."""
    from aslsd_b200 import simulate
    times = simulate.simulate_mhp(model, mu, kp, T_f, seed=1234)
    t0 = time.perf_counter()
    path = ref.ProcessPath(times, T_f)
    return path, time.perf_counter() - t0


def main():
    which = set(sys.argv[1:]) or {'B1', 'B3', 'B4'}
    ref = ref_shim.import_reference()
    warnings.simplefilter('ignore')
    if not hasattr(ref, 'ProcessPath'):
        from aslsd.stats.events.process_path import ProcessPath
        ref.ProcessPath = ProcessPath
    C = configs()
    if 'B4' in which or 'B3' in which:
        for name in ('C2', 'C3', 'C4'):
            model, mu, kp, T_f = C[name]
            d = model['d']
            path, dt_maps = reference_path(ref, model, mu, kp, T_f)
            n_events = int(sum(path.n_events))
            if 'B4' in which:
                emit(row='B4', what='ProcessPath.__init__: get_kappa_varpi + get_lag_sizes (time_ordering.py:9-119)',
                     config=name, n_events=n_events, seconds=dt_maps)
            if 'B3' in which:
                mhp = ref_shim.build_reference_model(ref, model)
                xs = x_at(mu, kp, d)
                rng = np.random.default_rng(1)
                ests = []
                for k in range(d):
                    est = ref.MHPStratEstim()
                    est.initialize(k, 50, mhp, path, **STRAT_KW)
                    ests.append(est)
                n_iter = 6
                for k in range(d):     # warm-up
                    ests[k].lse_k_grad_estimate(xs[k], rng=rng, grad_alloc=False)
                t0 = time.perf_counter()
                for _ in range(n_iter):
                    for k in range(d):
                        ests[k].lse_k_grad_estimate(xs[k], rng=rng, grad_alloc=False)
                dt = (time.perf_counter() - t0) / n_iter
                emit(row='B3', what='MHPStratEstim.lse_k_grad_estimate for every k (one ASLSD epoch without the solver step)',
                     config=name, n_events=n_events, s_per_epoch=dt, epochs_timed=n_iter)
    if 'B4x' in which:
        # two of the ten event types of C5 at full length (~10^6 events each)
        from aslsd_b200 import simulate
        model, times, T_f, xs = simulate.config_c5(d=10, n_basis=10, T_f=1.1e6, seed=1234)
        t0 = time.perf_counter()
        path = ref.ProcessPath(times[:2], T_f)
        emit(row='B4', what='ProcessPath.__init__: get_kappa_varpi + get_lag_sizes (time_ordering.py:9-119)',
             config='C5, 2 of the 10 event types', n_events=int(sum(path.n_events)),
             seconds=time.perf_counter() - t0,
             note='all 100 ordered pairs of C5 cost ~45x this (90 cross pairs instead of 2)')
    if 'B1' in which:
        # C1 in full
        model, mu, kp, T_f = C['C1']
        path, _ = reference_path(ref, model, mu, kp, T_f)
        mhp = ref_shim.build_reference_model(ref, model)
        est = ref.MHPExactEstim()
        est.initialize(0, 3, mhp, path)
        x = x_at(mu, kp, 1)[0]
        t0 = time.perf_counter()
        est.lse_k_grad_estimate(x)
        dt = time.perf_counter() - t0
        n = int(path.n_events[0])
        pairs = 2 * (n * (n - 1) // 2)           # phi_0 and upsilon_0_0, untruncated
        emit(row='B1', what='MHPExactEstim.lse_k_grad_estimate (literal O(N^2))', config='C1', n_events=n,
             seconds=dt, untruncated_pair_terms=pairs, pair_terms_per_s=pairs / dt)
        # C5 on a prefix: 10 types x ~400 events, sub-problem k = 0
        from aslsd_b200 import simulate
        model, times, T_f, xs = simulate.config_c5(T_f=400.)
        path = ref.ProcessPath(times, T_f)
        mhp = ref_shim.build_reference_model(ref, model)
        est = ref.MHPExactEstim()
        est.initialize(0, 3, mhp, path)
        t0 = time.perf_counter()
        est.lse_k_grad_estimate(np.array(xs[0], dtype=object))
        dt = time.perf_counter() - t0
        ns = [int(v) for v in path.n_events]
        pairs = 0
        for i in range(10):
            for j in range(10):
                pairs += int(np.sum(np.asarray(path.lag_sizes[i][j], dtype=np.int64)))   # upsilon_i_j
            pairs += int(np.sum(np.asarray(path.lag_sizes[0][i], dtype=np.int64)))       # phi_i (targets of type 0)
        emit(row='B1', what='MHPExactEstim.lse_k_grad_estimate, sub-problem k=0 of C5 on a prefix (T_f = 400)',
             config='C5 prefix', n_events=sum(ns), seconds=dt, untruncated_pair_terms=pairs,
             pair_terms_per_s=pairs / dt,
             note='untruncated pairs: the reference evaluates every earlier event; the device and the oracle '
                  'port count in-window pairs only')


if __name__ == '__main__':
    main()
