"""Import the read-only reference (saadlabyad/aslsd) in the build container.

The reference imports itself as the package ``aslsd`` and imports matplotlib
at module top (models/hawkes/linear/mhp.py:8, utilities/graphic_tools.py:6-7),
which is not installed here.  This shim puts a symlink
``<tmp>/aslsd -> /root/reference`` on ``sys.path`` and registers empty
matplotlib stub modules.  It is used only by the golden-vector generator and by
tests that are skipped when /root/reference is absent (e.g. on the GPU box).
"""
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get('ASLSD_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'optimize'))


_shim_dir = None


def make_importable():
    """Put ``aslsd`` (the reference) on sys.path and stub matplotlib; no import
    happens here.  Called from tests/conftest.py before ``aslsd_b200`` is first
    imported, so that the B200 estimators can derive from the reference's ABC."""
    global _shim_dir
    if _shim_dir is not None or not reference_available():
        return
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.animation',
                 'matplotlib.ticker'):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
    _shim_dir = tempfile.mkdtemp(prefix='aslsd_ref_')
    os.symlink(REFERENCE_ROOT, os.path.join(_shim_dir, 'aslsd'))
    sys.path.insert(0, _shim_dir)


def import_reference():
    """Returns a namespace with the reference classes used by the tests."""
    if not reference_available():
        raise ImportError('reference tree not present at %s' % REFERENCE_ROOT)
    make_importable()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from aslsd.models.hawkes.linear.mhp import MHP
        from aslsd.functionals.kernels.kernel import KernelModel
        from aslsd.functionals.kernels.basis_kernels.basis_kernel_exponential \
            import ExponentialKernel
        from aslsd.functionals.kernels.basis_kernels.basis_kernel_gaussian \
            import GaussianKernel
        from aslsd.stats.events.process_path import ProcessPath
        from aslsd.stats.events import time_ordering
        from aslsd.optimize.estimators.mhp_exact_estimator import MHPExactEstim
        from aslsd.optimize.estimators.mhp_stratified_estimator import \
            MHPStratEstim
        from aslsd.optimize.stratifications.general_stratification import \
            GeneralStratification
        from aslsd.optimize.stratifications.double_ix_stratification import \
            DoubleIxStratification
        from aslsd.utilities import useful_statistics
    ns = types.SimpleNamespace(
        MHP=MHP, KernelModel=KernelModel, ExponentialKernel=ExponentialKernel,
        GaussianKernel=GaussianKernel, ProcessPath=ProcessPath,
        time_ordering=time_ordering, MHPExactEstim=MHPExactEstim,
        MHPStratEstim=MHPStratEstim,
        GeneralStratification=GeneralStratification,
        DoubleIxStratification=DoubleIxStratification,
        useful_statistics=useful_statistics)
    return ns


def build_reference_model(ref, model):
    """Plain-data model description (oracle/exact.py) -> reference MHP."""
    def basis(b):
        cls = {'exp': ref.ExponentialKernel, 'gauss': ref.GaussianKernel}[b['type']]
        fixed = b.get('fixed', {})
        if not fixed:
            return cls()
        ixs = sorted(fixed)
        return cls(fixed_indices=list(ixs), fixed_vars=[fixed[i] for i in ixs])
    d = model['d']
    mat = [[ref.KernelModel([basis(b) for b in model['kernels'][k][i]])
            for i in range(d)] for k in range(d)]
    return ref.MHP(mat)
