"""Small device workloads for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_device.py

(compute-sanitizer is closed on the GPU pool this was developed on - profiles/r2_compute_sanitizer_closed.log -
so the script carries its own check for races: every call is repeated and must return the same bits;
the kernels have no atomics and a fixed summation order, so any difference is a race.)

Every kernel family of the library runs once on inputs small enough for the sanitizer's
slow-down: the lag-moment kernels (a 10-type model with 64 bins per stream: the warp-shaped kernel;
exponential / mixed models with more bins: the sorting kernel in its 512-thread shape), the per-pair kernels (targets near T_f), the sampled
sums, the point sums, the index maps and the simulator.  The sums are compared with the oracle, so
a race that changes a result fails the run as well."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402

import aslsd_b200  # noqa: E402
from aslsd_b200 import simulate  # noqa: E402


def main():
    import oracle_engine
    from cases import CASES, case_x
    eng = aslsd_b200.Engine(0)
    # ---- C5's model on a short stream: moment kernel (warp-shaped: streams of 64 bins), quadrature rows, boundary targets
    model, times, T_f, xs = simulate.config_c5(T_f=1500.)
    eng.set_events(times, T_f)
    eng.set_model(model)
    eng.use_weight_cache(False)
    d = model['d']
    out = eng.exact_sums_batch(list(range(d)), xs)
    ora = oracle_engine.OracleEngine()
    ora.set_events(times, T_f)
    ora.desc = eng.desc
    ora.plain = eng.desc.to_plain()
    ref = ora.exact_sums_packed(0, xs[0])
    err = np.max(np.abs(out[0] - ref) / (np.abs(ref) + 1e-300))
    print('c5 short stream: k = 0 max rel', err, eng.last_work_counts())
    assert err < 1e-10
    for rep in range(4):
        again = eng.exact_sums_batch(list(range(d)), xs)
        assert all(np.array_equal(a, b) for a, b in zip(out, again)), 'batched sums differ between runs'
    kappa, varpi, lag_sizes = eng.index_maps()
    assert len(kappa) == d
    # ---- smaller models: every basis family, the 512 x 1 shape (more than 64 bins), sampled and point sums
    for name in ('exp2d', 'gauss2d', 'mix2d', 'exp3d'):
        model, mu, kp, T_f = CASES[name]
        T_f = 40. * T_f
        times = eng.simulate(model, mu, kp, T_f, seed=5)
        eng.set_events(times, T_f)
        eng.set_model(model)
        ora.set_events(times, T_f)
        ora.desc = eng.desc
        ora.plain = eng.desc.to_plain()
        for k in range(model['d']):
            x = case_x(name)[k]
            got = eng.exact_sums_packed(k, x)
            ref = ora.exact_sums_packed(k, x)
            err = np.max(np.abs(got - ref) / (np.abs(ref) + 1e-300))
            print(name, k, 'exact max rel', err, eng.last_work_counts())
            assert err < 1e-9
            for rep in range(3):
                assert np.array_equal(got, eng.exact_sums_packed(k, x)), 'sums differ between runs'
    print('sanitize workloads ok')


if __name__ == '__main__':
    main()
