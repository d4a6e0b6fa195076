"""Independent LSE values for exponential models: the reference's O(N) Markov recursion.

Run in the build container only (needs /root/reference).  ``RecurrentExponential.get_lse_k``
(models/hawkes/linear/rec_exp.py:1692) computes the exact LSE_k of an exponential MHP by a
recursion over the events - a different algorithm from the pairwise sums of MHPExactEstim, hence
an independent check of the value at a scale where the literal pairwise reference takes minutes
(SURVEY 8c item 3).  It wants void marks on the path (rec_exp.py:1572).  Stored: the simulated
event times, the evaluation points and LSE_k for every k.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (HERE, os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE))):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from cases import CASES, case_x  # noqa: E402

RECEXP_CASES = {'recexp_exp1d': ('exp1d', 6000.), 'recexp_exp2d': ('exp2d', 1500.)}


def make_recexp(ref):
    import ref_shim
    from aslsd.functionals.kernels.basis_kernels.basis_kernel_exponential import ExponentialKernel
    from aslsd.functionals.kernels.kernel import KernelModel
    from aslsd.models.hawkes.linear.rec_exp import RecurrentExponential
    from aslsd.stats.events.process_path import ProcessPath
    for name, (case, T_f) in RECEXP_CASES.items():
        model, mu, kp, _ = CASES[case]
        d = model['d']
        mhp = ref_shim.build_reference_model(ref, model)
        arr = np.empty((d, d), dtype=object)
        for a in range(d):
            for b in range(d):
                arr[a][b] = np.asarray(kp[a][b], dtype=float)
        path = mhp.simulate(T_f, mu=np.asarray(mu, float), kernel_param=arr, seed=1234, verbose=False)
        rec = RecurrentExponential([[KernelModel([ExponentialKernel() for _ in model['kernels'][i][j]])
                                     for j in range(d)] for i in range(d)])
        void = ProcessPath(path.list_times, path.T_f,
                           list_marks=[np.zeros((len(t), 0)) for t in path.list_times])
        out = {'T_f': np.float64(T_f), 'd': np.int64(d), 'case': np.array(case)}
        for i in range(d):
            out['times/%d' % i] = np.asarray(path.list_times[i], dtype=float)
        for k in range(d):
            x = np.array(case_x(case)[k], dtype=float)
            out['k%d/x' % k] = x
            out['k%d/lse' % k] = np.float64(rec.get_lse_k(k, void, x_k=x, initialize=True))
        np.savez_compressed(os.path.join(HERE, '%s.npz' % name), **out)
        print('recexp', name, [len(t) for t in path.list_times], [float(out['k%d/lse' % k]) for k in range(d)])


if __name__ == '__main__':
    import ref_shim
    make_recexp(ref_shim.import_reference())
