"""Golden vectors of the benchmark's own model (config C5 of SURVEY.md 8d: ten event types, every
kernel a sum of ten Gaussians with beta and delta fixed) from the live reference, on a short path.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_c5model.py

The stream is the first T_f time units of the benchmark's stream generator
(``aslsd_b200.simulate.config_c5``), the evaluation points are its perturbed truth; the reference's
``MHPExactEstim.compute_objective`` gives the per-term sums and derivatives of sub-problems 0, 4 and 9
(one of them is the last type: the reference's late-bound upsilon closures, SURVEY quirk 4, all end up
on its kernels).  Pins oracle/exact.py at d = 10; the device is compared with the oracle on the same
model in tests/test_gpu_exact.py::test_c5_model_matches_oracle_every_k."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402
from cases import flatten_sums  # noqa: E402

T_F = 40.
KS = (0, 4, 9)


def main():
    ref = ref_shim.import_reference()
    from aslsd_b200 import simulate
    model, times, T_f, xs = simulate.config_c5(T_f=T_F)
    d = model['d']
    mhp = ref_shim.build_reference_model(ref, model)
    path = ref.ProcessPath([np.asarray(t, float) for t in times], T_f)
    out = {'T_f': np.float64(T_f), 'd': np.int64(d), 'ks': np.asarray(KS, np.int64)}
    for i in range(d):
        out['times/%d' % i] = np.asarray(times[i], dtype=float)
    for k in KS:
        t0 = time.time()
        est = ref.MHPExactEstim()
        est.initialize(k, 10, mhp, path)
        est.compute_objective(np.array(xs[k], dtype=object), compute_f_sum=True, compute_diff_f_sum=True)
        out['k%d/x' % k] = xs[k]
        out.update(flatten_sums(est.sum_f, est.sum_diff_f, 'k%d' % k))
        out['k%d/lse' % k] = np.float64(est.recombine_lse_k())
        est.recombine_diff_lse_theta_kpr = est.recombine_diff_lse_theta_kpr_kernel
        out['k%d/grad' % k] = np.asarray(ref.MHPStratEstim.recombine_grad_lse_k(est), dtype=float)
        print('k = %d: %.1f s, %d events' % (k, time.time() - t0, sum(len(t) for t in times)))
    np.savez_compressed(os.path.join(HERE, 'exact_c5model.npz'), **out)


if __name__ == '__main__':
    main()
