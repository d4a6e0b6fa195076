"""Golden vectors of the marked model (MTLH) exact sums, from the live reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_mtlh.py

For every case of ``cases.MTLH_CASES`` the reference's ``MTLH`` model is built (Exponential /
Gaussian kernels, ``ConstantBaseline``, ``LinearImpact`` / ``ConstantImpact``), a ``ProcessPath``
with marks is made from the seeded inputs of ``cases.mtlh_inputs`` and
``MTLHExactEstim.compute_objective(f + derivatives)`` is evaluated; M term, every per-term sum
(value, kernel / baseline / impact derivative blocks), the LSE and its gradient are stored.
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
from cases import MTLH_CASES, mtlh_inputs  # noqa: E402
from oracle import mtlh_exact as OM  # noqa: E402  (only for its flatten helper)


def build_reference_mtlh(model):
    ref_shim.make_importable()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from aslsd.models.hawkes.linear.mtlh import MTLH
        from aslsd.functionals.kernels.kernel import KernelModel
        from aslsd.functionals.kernels.basis_kernels.basis_kernel_exponential import ExponentialKernel
        from aslsd.functionals.kernels.basis_kernels.basis_kernel_gaussian import GaussianKernel
        from aslsd.functionals.baselines.baseline import BaselineModel
        from aslsd.functionals.baselines.basis_baselines.basis_baseline_constant import ConstantBaseline
        from aslsd.functionals.baselines.basis_baselines.basis_baseline_piececonst import PieceConstBaseline
        from aslsd.functionals.impact_functions.impact_function import ImpactFunction
        from aslsd.functionals.impact_functions.basis_impacts.basis_impact_linear import LinearImpact
        from aslsd.functionals.impact_functions.basis_impacts.basis_impact_constant import ConstantImpact
    d = model['d']

    def basis(b):
        cls = ExponentialKernel if b['type'] == 'exp' else GaussianKernel
        fixed = b.get('fixed') or {}
        if fixed:
            ix = sorted(fixed)
            return cls(fixed_indices=ix, fixed_vars=[fixed[v] for v in ix])
        return cls()

    def imp(spec):
        if spec['type'] == 'linear':
            return ImpactFunction([LinearImpact(mark_dim=spec.get('mark_dim', 1))])
        return ImpactFunction([ConstantImpact()])
    km = [[KernelModel([basis(b) for b in model['kernels'][k][i]]) for i in range(d)] for k in range(d)]
    specs = model.get('baselines', [{'type': 'constant'}] * d)
    bl = [BaselineModel([ConstantBaseline() if b['type'] == 'constant'
                         else PieceConstBaseline(beta=np.asarray(b['beta'], dtype=float))]) for b in specs]
    im = [[imp(model['impacts'][k][i]) for i in range(d)] for k in range(d)]
    gauss = any(b['type'] == 'gauss' for row in model['kernels'] for ker in row for b in ker)
    # GaussianKernel.make_K is not implemented upstream (basis_kernel_gaussian.py:542-546), nor is
    # any kernel-side K for a PieceConstBaseline (KeyError in the kernels' dict_K): the baseline-side
    # K is the only route then (mtlh.py:1055-1064)
    return MTLH(km, _baselines_vec=bl, _impact_matrix=im,
                K_from_baseline=gauss or any(b['type'] != 'constant' for b in specs))


def main():
    ref_shim.make_importable()
    from aslsd.stats.events.process_path import ProcessPath
    from aslsd.optimize.estimators.mtlh_exact_estimator import MTLHExactEstim
    for name, (model, xs, n_events, T_f) in MTLH_CASES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        d = model['d']
        mtlh = build_reference_mtlh(model)
        times, marks, T_f = mtlh_inputs(name)
        path = ProcessPath(times, T_f, list_marks=marks)
        out = {'T_f': np.float64(T_f), 'd': np.int64(d)}
        for i in range(d):
            out['times/%d' % i] = times[i]
            out['marks/%d' % i] = marks[i]
        for k in range(d):
            est = MTLHExactEstim()
            est.initialize(k, 10, mtlh, path)
            x = np.asarray(xs[k], dtype=float)
            assert len(x) == mtlh.n_param_k[k], (name, len(x), mtlh.n_param_k[k])
            est.compute_objective(x, compute_f_sum=True, compute_diff_f_sum=True)
            out['k%d/x' % k] = x
            out.update(OM.flatten(est.M_term, est.diff_M_term, est.sum_f, est.sum_diff_f, 'k%d' % k))
            out['k%d/lse' % k] = np.float64(est.recombine_lse_k())
            try:
                out['k%d/grad' % k] = np.asarray(est.recombine_grad_lse_k(), dtype=float)
            except TypeError:
                # ConstantImpact has no parameters and the reference's gradient assembly indexes
                # interval_map_imp[k][p], which is None then (mtlh_exact_estimator.py:664): the
                # sums are pinned, the gradient is not available from the reference
                pass
        np.savez_compressed(os.path.join(HERE, '%s.npz' % name), **out)
        print(name, n_events, 'lse', [float(out['k%d/lse' % k]) for k in range(d)])


if __name__ == '__main__':
    main()
