"""Host-side model description for the B200 engine.

Lightweight mirrors of the reference's model objects - only what the LSE path
reads from them - plus ``describe_model`` which extracts the device descriptor
from *either* these mirrors, the reference's own ``MHP`` object (duck typing on
``_kernel_matrix[k][i]._basis_kernels[r]``: class name, ``fixed_indices``,
``fixed_vars``) or a plain dict ``{'d', 'kernels'}``.

Reference interfaces mirrored:
  * BasisKernel fixed variables   functionals/kernels/basis_kernel.py:113-130, :364-401
  * KernelModel parameter maps    functionals/kernels/kernel.py:97-117
  * MHP parameter maps            models/hawkes/linear/mhp.py:276-299
  * MHP.init_estimator            models/hawkes/linear/mhp.py:1525-1538
"""
import numpy as np

from aslsd_b200 import capi

_N_VARS = {'ExponentialKernel': 2, 'GaussianKernel': 3}
# get_var_lower_bounds of basis_kernel_exponential.py:366-367 / basis_kernel_gaussian.py:426-427
_LOWER_BOUNDS = {'ExponentialKernel': [1e-10, 1e-10], 'GaussianKernel': [1e-8, 1e-8, 0.]}
_TYPE_ID = {'ExponentialKernel': capi.BASIS_EXPONENTIAL,
            'GaussianKernel': capi.BASIS_GAUSSIAN}
_SHORT = {'exp': 'ExponentialKernel', 'gauss': 'GaussianKernel'}


class BasisKernel:
    """A basis kernel with optionally fixed variables
    (basis_kernel.py:113-130)."""
    n_vars = None

    def __init__(self, fixed_indices=None, fixed_vars=None):
        if fixed_indices is None:
            fixed_indices, fixed_vars = [], []
        fixed_indices = np.atleast_1d(np.asarray(fixed_indices, dtype=int))
        fixed_vars = np.atleast_1d(np.asarray(fixed_vars, dtype=float))
        if len(fixed_indices) != len(fixed_vars):
            raise ValueError('fixed_indices and fixed_vars must have the same length')
        order = np.argsort(fixed_indices, kind='stable')
        self.fixed_indices = fixed_indices[order]
        self.fixed_vars = fixed_vars[order]
        self.n_fixed_vars = len(self.fixed_vars)

    def get_n_vars(self):
        return self.n_vars

    def get_n_param(self):
        return self.n_vars - self.n_fixed_vars


class ExponentialKernel(BasisKernel):
    """phi(t) = omega beta exp(-beta t) (basis_kernel_exponential.py:387-388)."""
    n_vars = 2


class GaussianKernel(BasisKernel):
    """phi(t) = omega N(t; delta, beta) (basis_kernel_gaussian.py:448-452)."""
    n_vars = 3


class KernelModel:
    """Sum of basis kernels (kernel.py:14-92)."""

    def __init__(self, _basis_kernels):
        if isinstance(_basis_kernels, BasisKernel):
            _basis_kernels = [_basis_kernels]
        self._basis_kernels = list(_basis_kernels)
        self.n_basis_ker = len(self._basis_kernels)
        self.vec_n_param = [b.get_n_param() for b in self._basis_kernels]
        self.n_param = sum(self.vec_n_param)


class MHP:
    """Multivariate Hawkes model: only the parameter book-keeping of
    models/hawkes/linear/mhp.py:216-299."""

    def __init__(self, _kernel_matrix):
        if isinstance(_kernel_matrix, dict):
            # plain description {'d', 'kernels'}: build the mirror objects
            _kernel_matrix = [[KernelModel([_basis_from_plain(b) for b in ker]) for ker in row]
                              for row in _kernel_matrix['kernels']]
        self._kernel_matrix = [list(row) for row in _kernel_matrix]
        self.d = len(self._kernel_matrix)
        self.desc = describe_model(self)
        self.matrix_n_param = self.desc.matrix_n_param
        self.ix_map = self.desc.ix_map
        self.interval_map = self.desc.interval_map
        # value closures in the reference's layout (mhp.py:475-540): the stratified estimator takes
        # the saturated value of its single-index terms from them (aslsd_b200.saturation)
        from aslsd_b200.saturation import value_closures
        self.psi, self.upsilon = value_closures(self.desc)

    def init_estimator(self, estimator, k):
        estimator.n_param_k = 1 + sum(self.matrix_n_param[k])
        estimator.matrix_n_param = self.matrix_n_param
        estimator.ix_map = self.ix_map
        estimator.interval_map = self.interval_map

    def lower_bounds(self, k):
        """Lower bounds of x_k: 1e-10 for mu_k (mhp.py:1701), then the free
        variables' bounds of every basis kernel."""
        out = [1e-10]
        for i in range(self.d):
            for type_name, free_mask, _ in self.desc.bases[k][i]:
                out += [b for v, b in enumerate(_LOWER_BOUNDS[type_name]) if (free_mask >> v) & 1]
        return np.array(out)

    def fit(self, process_path, x_0, n_iter=1000, solvers=None, solver_args=None,
            exact_grad=False, estimators=None, is_log_lse=False, is_grad_target=False,
            is_log_ixs=False, is_log_allocs=False, is_log_total_estimates=False,
            is_log_strata_estimates=False, n_exact_single=None,
            n_samples_adaptive_single=None, nonadaptive_sample_size_single=None,
            single_strfs=None, n_samples_adaptive_double=None,
            nonadaptive_sample_size_double=None, double_strfs=None, rng=None, seed=1234,
            engine=None, callback=None):
        """The estimator / solver loop of MHP.fit (mhp.py:1609-1804) for an
        explicit starting point ``x_0`` (list of d parameter vectors): for each
        dimension k, n_iter[k] iterations of gradient estimate -> solver step ->
        projection on the parameter bounds, all k sharing one RNG.  Returns
        ``(x, log)`` with ``log['param'][k]`` / ``log['grad'][k]`` the iterates."""
        from aslsd_b200.estimators import MHPExactEstimB200, MHPStratEstimB200
        from aslsd_b200.solvers import ADAM
        d = self.d
        if rng is None:
            rng = np.random.default_rng(seed)
        if not isinstance(n_iter, (list, np.ndarray)):
            n_iter = [n_iter] * d
        if estimators is None:
            cls = MHPExactEstimB200 if exact_grad else MHPStratEstimB200
            estimators = [cls(is_grad_target=is_grad_target, is_log_ixs=is_log_ixs,
                              is_log_allocs=is_log_allocs,
                              is_log_total_estimates=is_log_total_estimates,
                              is_log_strata_estimates=is_log_strata_estimates,
                              is_log_lse=is_log_lse, engine=engine) for _ in range(d)]
        for k in range(d):
            estimators[k].initialize(
                k, n_iter[k], self, process_path, n_exact_single=n_exact_single,
                n_samples_adaptive_single=n_samples_adaptive_single,
                nonadaptive_sample_size_single=nonadaptive_sample_size_single,
                single_strfs=single_strfs, n_samples_adaptive_double=n_samples_adaptive_double,
                nonadaptive_sample_size_double=nonadaptive_sample_size_double,
                double_strfs=double_strfs)
        if solvers is None:
            solvers = [ADAM(**(solver_args or {})) for _ in range(d)]
        log = {'param': [], 'grad': []}
        x = [None] * d
        for k in range(d):
            x_k = np.asarray(x_0[k], dtype=np.float64)
            lower = self.lower_bounds(k)
            params = np.zeros((n_iter[k] + 1, len(x_k)))
            grads = np.zeros((n_iter[k], len(x_k)))
            params[0] = x_k
            for t in range(n_iter[k]):
                g_t = estimators[k].lse_k_grad_estimate(x_k, rng=rng, grad_alloc=is_grad_target)
                grads[t] = g_t
                x_k = np.clip(solvers[k].iterate(t, x_k, g_t), lower, np.inf)
                params[t + 1] = x_k
                if callback is not None:
                    callback(k, t, x_k, g_t)
            log['param'].append(params)
            log['grad'].append(grads)
            x[k] = x_k
        self.fit_estim = estimators
        self.fitted_x = x
        return x, log


    def fit_episodes(self, episodes, x_0, n_iter=1000, solvers=None, solver_args=None,
                     exact_grad=False, estimators=None, is_log_lse=False, is_grad_target=False,
                     is_log_ixs=False, is_log_allocs=False, is_log_total_estimates=False,
                     is_log_strata_estimates=False, n_exact_single=None,
                     n_samples_adaptive_single=None, nonadaptive_sample_size_single=None,
                     single_strfs=None, n_samples_adaptive_double=None,
                     nonadaptive_sample_size_double=None, double_strfs=None, rng=None, seed=1234,
                     process_group=None, callback=None):
        """The loop of MHP.fit_episodes (mhp.py:1806-2008): one estimator per (k, path), the
        gradient of an iteration is the mean of the per-path gradient estimates
        (``np.mean(g_per_path, axis=0)``, :1983-1989), then solver step and projection as in
        ``fit``.  Every path keeps its own device context (``events.engine_for``), so the events
        of all episodes stay resident.

        ``estimators``: None, one estimator (copied for every (k, path)) or a d x n_episodes
        nested list, like upstream.  ``process_group``: a ``torch.distributed`` group - the
        episodes are dealt round-robin to its ranks, each rank evaluates its own paths and one
        all-reduce per iteration sums the per-path gradients (the extra data-parallel axis of
        SURVEY 8f row 4); every rank returns the same iterates.

        Deviation from upstream, on purpose: the reference calls
        ``lse_k_grad_estimate(x_k, rng)`` (:1986), which binds the Generator to ``grad_alloc`` and
        leaves ``rng=None`` - the stratified estimator then draws from fresh OS entropy at every
        call (mhp_stratified_estimator.py:1409-1413) and a fit is not reproducible.  Here the
        Generator is passed as ``rng`` (and ``grad_alloc=True``, which is what a truthy Generator
        amounts to); with exact estimators the two coincide."""
        import copy
        from aslsd_b200.estimators import MHPExactEstimB200, MHPStratEstimB200
        from aslsd_b200.solvers import ADAM
        d = self.d
        n_episodes = episodes.n_episodes
        if n_episodes == 0:
            raise ValueError('fit_episodes: no episodes')
        if rng is None:
            rng = np.random.default_rng(seed)
        if not isinstance(n_iter, (list, np.ndarray)):
            n_iter = [n_iter] * d
        rank, world = 0, 1
        if process_group is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
        mine = [ix for ix in range(n_episodes) if ix % world == rank]
        if estimators is None:
            cls = MHPExactEstimB200 if exact_grad else MHPStratEstimB200
            estimators = [[cls(is_grad_target=is_grad_target, is_log_ixs=is_log_ixs,
                               is_log_allocs=is_log_allocs,
                               is_log_total_estimates=is_log_total_estimates,
                               is_log_strata_estimates=is_log_strata_estimates,
                               is_log_lse=is_log_lse) for _ in range(n_episodes)] for _ in range(d)]
        elif not isinstance(estimators, (list, tuple)):
            estimators = [[copy.deepcopy(estimators) for _ in range(n_episodes)] for _ in range(d)]
        for k in range(d):
            for ix in mine:
                estimators[k][ix].initialize(
                    k, n_iter[k], self, episodes.list_paths[ix], n_exact_single=n_exact_single,
                    n_samples_adaptive_single=n_samples_adaptive_single,
                    nonadaptive_sample_size_single=nonadaptive_sample_size_single,
                    single_strfs=single_strfs, n_samples_adaptive_double=n_samples_adaptive_double,
                    nonadaptive_sample_size_double=nonadaptive_sample_size_double,
                    double_strfs=double_strfs)
        if solvers is None:
            solvers = [ADAM(**(solver_args or {})) for _ in range(d)]
        elif not isinstance(solvers, (list, tuple)):
            solvers = [copy.deepcopy(solvers) for _ in range(d)]
        log = {'param': [], 'grad': []}
        x = [None] * d
        for k in range(d):
            x_k = np.asarray(x_0[k], dtype=np.float64)
            lower = self.lower_bounds(k)
            params = np.zeros((n_iter[k] + 1, len(x_k)))
            grads = np.zeros((n_iter[k], len(x_k)))
            params[0] = x_k
            for t in range(n_iter[k]):
                g_per_path = np.zeros((n_episodes, len(x_k)))
                for ix in mine:
                    g_per_path[ix] = estimators[k][ix].lse_k_grad_estimate(x_k, rng=rng, grad_alloc=True)
                if world > 1:
                    import torch
                    import torch.distributed as dist
                    buf = torch.from_numpy(g_per_path)
                    if dist.get_backend(process_group) == 'nccl':
                        dev = buf.cuda()
                        dist.all_reduce(dev, group=process_group)
                        g_per_path = dev.cpu().numpy()
                    else:
                        dist.all_reduce(buf, group=process_group)
                g_t = np.mean(g_per_path, axis=0)
                grads[t] = g_t
                x_k = np.clip(solvers[k].iterate(t, x_k, g_t), lower, np.inf)
                params[t + 1] = x_k
                if callback is not None:
                    callback(k, t, x_k, g_t)
            log['param'].append(params)
            log['grad'].append(grads)
            x[k] = x_k
        self.fit_estim = estimators
        self.fitted_x = x
        return x, log


def _basis_from_plain(b):
    cls = {'exp': ExponentialKernel, 'gauss': GaussianKernel}[b['type']]
    fixed = b.get('fixed', {})
    ixs = sorted(fixed)
    return cls(fixed_indices=ixs, fixed_vars=[fixed[v] for v in ixs])


class ModelDesc:
    """Device descriptor of a model + the reference's parameter maps."""

    def __init__(self, d, bases):
        # bases[k][i] = list of (type_name, free_mask, fixed[3])
        self.d = d
        self.bases = bases
        self.matrix_n_param = [[sum(bin(b[1]).count('1') for b in bases[k][i])
                                for i in range(d)] for k in range(d)]
        # mhp.py:276-299
        self.ix_map = []
        self.interval_map = []
        for k in range(d):
            row = [-1]
            imap = []
            pos = 1
            for i in range(d):
                n = self.matrix_n_param[k][i]
                imap.append([pos, pos + n])
                row.extend({'ker': i, 'par': r} for r in range(n))
                pos += n
            self.ix_map.append(row)
            self.interval_map.append(imap)

    def n_param_k(self, k):
        return 1 + sum(self.matrix_n_param[k])

    def c_arrays(self):
        d = self.d
        n_basis = (capi.ctypes.c_int32 * (d * d))()
        flat = [b for k in range(d) for i in range(d) for b in self.bases[k][i]]
        arr = (capi.Basis * len(flat))()
        for k in range(d):
            for i in range(d):
                n_basis[k * d + i] = len(self.bases[k][i])
        for ix, (type_name, free_mask, fixed) in enumerate(flat):
            arr[ix].type = _TYPE_ID[type_name]
            arr[ix].free_mask = free_mask
            for v in range(3):
                arr[ix].fixed[v] = fixed[v]
        return n_basis, arr

    def to_plain(self):
        """Plain-data form used by the tests' oracle."""
        short = {v: k for k, v in _SHORT.items()}
        out = []
        for k in range(self.d):
            row = []
            for i in range(self.d):
                ker = []
                for type_name, free_mask, fixed in self.bases[k][i]:
                    nv = _N_VARS[type_name]
                    ker.append({'type': short[type_name],
                                'fixed': {v: fixed[v] for v in range(nv)
                                          if not (free_mask >> v) & 1}})
                row.append(ker)
            out.append(row)
        return {'d': self.d, 'kernels': out}


def _describe_basis(b):
    if isinstance(b, dict):
        type_name = _SHORT[b['type']]
        fixed_map = dict(b.get('fixed', {}))
    else:
        type_name = type(b).__name__
        fixed_map = {int(ix): float(val)
                     for ix, val in zip(np.atleast_1d(b.fixed_indices),
                                        np.atleast_1d(b.fixed_vars))}
    if type_name not in _N_VARS:
        # mirrors the NotImplementedError of mhp.py:538-540 for unsupported kernels
        raise NotImplementedError(
            'basis kernel %s is outside the B200 engine scope '
            '(ExponentialKernel, GaussianKernel)' % type_name)
    nv = _N_VARS[type_name]
    free_mask = 0
    fixed = [0., 0., 0.]
    for v in range(nv):
        if v in fixed_map:
            fixed[v] = float(fixed_map[v])
        else:
            free_mask |= 1 << v
    return (type_name, free_mask, fixed)


def describe_model(model):
    """Build the ModelDesc of a reference MHP, a mirror MHP, or a plain dict."""
    if isinstance(model, ModelDesc):
        return model
    if isinstance(model, dict):
        d = model['d']
        bases = [[[_describe_basis(b) for b in model['kernels'][k][i]]
                  for i in range(d)] for k in range(d)]
        return ModelDesc(d, bases)
    mat = getattr(model, '_kernel_matrix', None)
    if mat is None:
        raise TypeError('cannot describe model of type %s' % type(model).__name__)
    d = len(mat)
    bases = [[[_describe_basis(b) for b in mat[k][i]._basis_kernels]
              for i in range(d)] for k in range(d)]
    return ModelDesc(d, bases)
