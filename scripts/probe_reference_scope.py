"""Which basis kernels and baselines outside the engine's scope can the REFERENCE's own LSE estimators
evaluate at all?  Run in the build container (needs /root/reference):

    python scripts/probe_reference_scope.py > profiles/r2_reference_scope_probe.log

Every candidate is put into the smallest model that holds it (one event type, one basis) and one
gradient estimate is asked of MHPExactEstim / MTLHExactEstim; the exception, if any, is the
reference's own.  SURVEY.md section 8 rows f-1 (non-constant baselines) and f-3 (remaining basis
kernels) are scoped by this list."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import ref_shim  # noqa: E402

ref_shim.make_importable()
warnings.simplefilter('ignore')
from aslsd.models.hawkes.linear.mhp import MHP  # noqa: E402
from aslsd.models.hawkes.linear.mtlh import MTLH  # noqa: E402
from aslsd.functionals.kernels.kernel import KernelModel  # noqa: E402
from aslsd.functionals.kernels.basis_kernels.basis_kernel_exponential import ExponentialKernel  # noqa: E402
from aslsd.functionals.kernels.basis_kernels.basis_kernel_gaussian import GaussianKernel  # noqa: E402
from aslsd.functionals.kernels.basis_kernels.basis_kernel_delayedexp import DelayedExponentialKernel  # noqa: E402
from aslsd.functionals.kernels.basis_kernels.basis_kernel_squaredpow import SquaredPowerLawKernel  # noqa: E402
from aslsd.functionals.kernels.basis_kernels.basis_kernel_triangle import TriangularKernel  # noqa: E402
from aslsd.functionals.kernels.basis_kernels.basis_kernel_rayleigh import RayleighKernel  # noqa: E402
from aslsd.functionals.kernels.basis_kernels.basis_kernel_null import NullKernel  # noqa: E402
from aslsd.functionals.baselines.baseline import BaselineModel  # noqa: E402
from aslsd.functionals.baselines.basis_baselines.basis_baseline_constant import ConstantBaseline  # noqa: E402
from aslsd.functionals.baselines.basis_baselines.basis_baseline_linear import LinearBaseline  # noqa: E402
from aslsd.functionals.baselines.basis_baselines.basis_baseline_piececonst import PieceConstBaseline  # noqa: E402
from aslsd.functionals.baselines.basis_baselines.basis_baseline_cosine import CosineBaseline  # noqa: E402
from aslsd.functionals.impact_functions.impact_function import ImpactFunction  # noqa: E402
from aslsd.functionals.impact_functions.basis_impacts.basis_impact_linear import LinearImpact  # noqa: E402
from aslsd.stats.events.process_path import ProcessPath  # noqa: E402
from aslsd.optimize.estimators.mhp_exact_estimator import MHPExactEstim  # noqa: E402
from aslsd.optimize.estimators.mhp_stratified_estimator import MHPStratEstim  # noqa: E402
from aslsd.optimize.estimators.mtlh_exact_estimator import MTLHExactEstim  # noqa: E402

rng = np.random.default_rng(1)
T_f = 50.
times = [np.sort(rng.uniform(0, T_f, 60)), np.sort(rng.uniform(0, T_f, 40))]


def report(label, fn):
    try:
        g = fn()
        ok = g is not None and np.all(np.isfinite(np.asarray(g, dtype=float)))
        print('%-58s %s' % (label, 'runs, gradient finite' if ok else 'runs, gradient not finite: %s' % (g,)))
    except Exception as e:      # the reference's own failure
        print('%-58s FAILS  %s: %s' % (label, type(e).__name__, ' '.join(str(e).split())[:110]))


def mhp_grad(matrix, est_cls=MHPExactEstim):
    def run():
        model = MHP(matrix)
        d = len(matrix)
        path = ProcessPath(times[:d], T_f)
        est = est_cls()
        est.initialize(0, 5, model, path)
        x = 0.5 + 0.1 * np.arange(1 + sum(model.matrix_n_param[0]))
        kw = {'rng': np.random.default_rng(2)} if est_cls is MHPStratEstim else {}
        return est.lse_k_grad_estimate(x, **kw)
    return run


print('== basis kernels through MHPExactEstim.lse_k_grad_estimate (one event type, one basis unless stated)')
report('ExponentialKernel', mhp_grad([[KernelModel(ExponentialKernel())]]))
report('GaussianKernel', mhp_grad([[KernelModel(GaussianKernel())]]))
report('DelayedExponentialKernel', mhp_grad([[KernelModel(DelayedExponentialKernel())]]))
report('DelayedExponentialKernel, delta fixed',
       mhp_grad([[KernelModel(DelayedExponentialKernel(fixed_indices=[2], fixed_vars=[0.5]))]]))
report('DelayedExponentialKernel, beta and delta fixed',
       mhp_grad([[KernelModel(DelayedExponentialKernel(fixed_indices=[1, 2], fixed_vars=[1.2, 0.5]))]]))
report('DelayedExponentialKernel (stratified estimator)',
       mhp_grad([[KernelModel(DelayedExponentialKernel(fixed_indices=[2], fixed_vars=[0.5]))]], MHPStratEstim))
report('TriangularKernel', mhp_grad([[KernelModel(TriangularKernel())]]))
report('RayleighKernel', mhp_grad([[KernelModel(RayleighKernel())]]))
report('NullKernel beside exponential kernels (2 types)',
       mhp_grad([[KernelModel(ExponentialKernel()), KernelModel(NullKernel())],
                 [KernelModel(ExponentialKernel()), KernelModel(ExponentialKernel())]]))
report('SquaredPowerLawKernel', mhp_grad([[KernelModel(SquaredPowerLawKernel())]]))
report('SquaredPowerLawKernel, two bases',
       mhp_grad([[KernelModel([SquaredPowerLawKernel(), SquaredPowerLawKernel()])]]))
report('SquaredPowerLawKernel, two event types',
       mhp_grad([[KernelModel(SquaredPowerLawKernel()), KernelModel(SquaredPowerLawKernel())],
                 [KernelModel(SquaredPowerLawKernel()), KernelModel(SquaredPowerLawKernel())]]))


def mtlh_grad(make_baseline, k_from_baseline):
    def run():
        model = MTLH([[KernelModel(ExponentialKernel())]], _baselines_vec=[BaselineModel([make_baseline()])],
                     _impact_matrix=[[ImpactFunction([LinearImpact(mark_dim=1)])]], K_from_baseline=k_from_baseline)
        path = ProcessPath(times[:1], T_f, list_marks=[np.random.default_rng(3).uniform(0., 1., size=(len(times[0]), 1))])
        est = MTLHExactEstim()
        est.initialize(0, 5, model, path)
        x = 0.5 + 0.1 * np.arange(model.n_param_k[0])
        return est.lse_k_grad_estimate(x)
    return run


print('== baselines through MTLHExactEstim.lse_k_grad_estimate (exponential kernel, linear impact)')
for kfb in (True, False):
    tag = ' (K from the baseline)' if kfb else ' (K from the kernel)'
    report('ConstantBaseline' + tag, mtlh_grad(ConstantBaseline, kfb))
    report('LinearBaseline' + tag, mtlh_grad(LinearBaseline, kfb))
    report('PieceConstBaseline' + tag, mtlh_grad(lambda: PieceConstBaseline(beta=np.array([0., 20., 35., T_f])), kfb))
    report('CosineBaseline' + tag, mtlh_grad(CosineBaseline, kfb))
