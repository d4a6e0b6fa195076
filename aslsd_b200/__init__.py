"""aslsd_b200: B200-native engine for the LSE hot path of saadlabyad/aslsd.

Host-side mirrors of the reference's interfaces for this path (same names,
argument meaning and error behaviour) over a C-ABI CUDA library
(include/aslsd_b200.h).  There is no CPU fallback.
"""
from aslsd_b200.engine import Engine, EngineError  # noqa: F401
from aslsd_b200.events import Episodes, ProcessPath, get_kappa_varpi, get_lag_sizes  # noqa: F401
from aslsd_b200.model import (MHP, KernelModel, ExponentialKernel,  # noqa: F401
                              GaussianKernel, describe_model)

from aslsd_b200.estimators import MHPExactEstimB200, MHPStratEstimB200  # noqa: F401
from aslsd_b200.mtlh import MTLH, MTLHExactEstimB200, MTLHStratEstimB200  # noqa: F401
from aslsd_b200.stratifications import (GeneralStratification,  # noqa: F401
                                        DoubleIxStratification)
from aslsd_b200.solvers import ADAM  # noqa: F401
from aslsd_b200.residuals import (get_intensity_at_jumps, get_residuals, get_residuals_k,  # noqa: F401
                                  get_residuals_k_mtlh, get_residuals_mtlh)

__all__ = ['get_intensity_at_jumps', 'get_residuals', 'get_residuals_k', 'get_residuals_k_mtlh', 'get_residuals_mtlh',
           'MHPExactEstimB200', 'MHPStratEstimB200', 'MTLH', 'MTLHExactEstimB200', 'MTLHStratEstimB200', 'GeneralStratification',
           'DoubleIxStratification', 'ADAM', 'Engine', 'EngineError', 'Episodes', 'ProcessPath', 'get_kappa_varpi', 'get_lag_sizes',
           'MHP', 'KernelModel', 'ExponentialKernel', 'GaussianKernel', 'describe_model']
