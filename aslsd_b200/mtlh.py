"""Marked, time-dependent-baseline linear Hawkes models (MTLH) on the B200: exact LSE sums.

``MTLHExactEstimB200`` exposes the interface ``MTLH.fit(..., estimators=...)`` drives
(models/hawkes/linear/mtlh.py:2368-2397) and leaves behind the state of the reference's
``MTLHExactEstim`` (optimize/estimators/mtlh_exact_estimator.py): ``M_term``, ``diff_M_term``,
``sum_f[log_name]`` and the nested ``sum_diff_f[log_name][ix_f]([ix_func])[param]`` blocks of
``make_null_sumarray`` (:303-319), for the terms ``mu``, ``phi_i``, ``upsilonzero_i``, ``K_i``,
``upsilon_i_j`` (:181-190).

Division of labour (BASELINE.json north_star: the functionals stay the reference's Python):
  * the impact functions I_ki(xi) and their parameter derivatives are evaluated per event on
    the host by the model's own closures (``model.impact`` / ``model.diff_impact``,
    functionals/impact_functions/impact_function.py:116-131) - O(N) numpy - and uploaded as
    per-event weight vectors (``aslsd_set_weights``);
  * the O(N x window) pair sums and the per-event sums run on the device with those weights
    (``aslsd_exact_sums_weighted``): phi-term x I_ki(xi^i_n) (:838), upsilon-term x
    I_ki(xi^i_m) I_kj(xi^j_n) (:896-897, :950), K_i x I_ki(xi^i_m) (:707-753),
    upsilonzero_i x I_ki(xi^i_m)^2 (:755-805).  The derivative blocks with respect to the
    impact parameters (:847-853, :907-918, :963-971) are the same value sums with one weight
    vector replaced by the derivative of the impact;
  * baselines: constant only (K = mu psi, M = mu^2: basis_baseline_constant.py:30-65); other
    baselines raise NotImplementedError, as do the sampled (stratified) estimators for marked
    models.

``MTLH`` below is the plain-data mirror of the reference model's parameter book-keeping
(mtlh.py:336-635, :983-1077, :1971-1996) used where the reference package is absent.
"""
import copy
import itertools

import numpy as np

from aslsd_b200.engine import term_layout, unpack_terms
from aslsd_b200.estimators import _B200Estimator
from aslsd_b200.model import KernelModel, _basis_from_plain, describe_model


def _impact_n_param(spec):
    return int(spec.get('mark_dim', 1)) if spec['type'] == 'linear' else 0


class MTLH:
    """Mirror of the reference's MTLH for constant baselines and Linear / Constant impacts
    (basis_impact_linear.py:36-52, basis_impact_constant.py:33-40)."""

    def __init__(self, model):
        self.d = d = model['d']
        self._kernel_matrix = [[KernelModel([_basis_from_plain(b) for b in ker]) for ker in row]
                               for row in model['kernels']]
        self.impacts = [[dict(spec) for spec in row] for row in model['impacts']]
        self.desc = describe_model(self)
        self.vector_n_param_mu = [1] * d
        self.matrix_n_param_ker = self.desc.matrix_n_param
        self.matrix_n_param_imp = [[_impact_n_param(self.impacts[k][i]) for i in range(d)]
                                   for k in range(d)]
        self.interval_map_ker = [[None] * d for _ in range(d)]
        self.interval_map_imp = [[None] * d for _ in range(d)]
        self.ix_map_ker = [dict() for _ in range(d)]
        self.ix_map_imp = [dict() for _ in range(d)]
        self.n_param_k = [0] * d
        for k in range(d):
            pos = 1
            for i in range(d):
                n = self.matrix_n_param_ker[k][i]
                self.interval_map_ker[k][i] = [pos, pos + n]
                for r in range(n):
                    self.ix_map_ker[k][pos + r] = {'ker': i, 'par': r}
                pos += n
            for i in range(d):
                n = self.matrix_n_param_imp[k][i]
                self.interval_map_imp[k][i] = [pos, pos + n]
                for r in range(n):
                    self.ix_map_imp[k][pos + r] = {'imp': i, 'par': r}
                pos += n
            self.n_param_k[k] = pos
        # constant baseline (basis_baseline_constant.py:30-49)
        self.mu = [lambda t, x_mu: x_mu[0] * np.ones(len(t)) for _ in range(d)]
        self.diff_mu = [lambda t, ix, x_mu: np.ones(len(t)) for _ in range(d)]
        self.M = [lambda t, x_mu: x_mu[0] ** 2 for _ in range(d)]
        self.diff_M = [lambda t, ix, x_mu: 2. * x_mu[0] for _ in range(d)]
        self.impact = [[self._make_impact(self.impacts[k][i]) for i in range(d)] for k in range(d)]
        self.diff_impact = [[self._make_diff_impact(self.impacts[k][i]) for i in range(d)]
                            for k in range(d)]

    @staticmethod
    def _make_impact(spec):
        if spec['type'] == 'linear':
            return lambda xi, params: 1. + np.sum(xi * np.asarray(params, dtype=float), axis=xi.ndim - 1)
        return lambda xi, params: np.ones(len(xi))

    @staticmethod
    def _make_diff_impact(spec):
        return lambda xi, ix_diff, params: xi[:, ix_diff]

    def xk2matrix_params(self, k, x_k):
        """mtlh.py:595-635"""
        d = self.d
        x_mu = np.array([x_k[ix] for ix in range(self.vector_n_param_mu[k])], dtype=float)
        x_ker = [np.array(x_k[self.interval_map_ker[k][i][0]:self.interval_map_ker[k][i][1]], dtype=float)
                 for i in range(d)]
        x_imp = [np.array(x_k[self.interval_map_imp[k][i][0]:self.interval_map_imp[k][i][1]], dtype=float)
                 for i in range(d)]
        return x_mu, x_ker, x_imp

    def init_estimator(self, estimator, k):
        """mtlh.py:1971-1996 (the functionals the device replaces are not attached)."""
        estimator.n_param_k = self.n_param_k[k]
        for name in ('vector_n_param_mu', 'matrix_n_param_ker', 'matrix_n_param_imp', 'ix_map_ker',
                     'ix_map_imp', 'interval_map_ker', 'interval_map_imp', 'xk2matrix_params', 'M',
                     'diff_M', 'mu', 'diff_mu', 'impact', 'diff_impact'):
            setattr(estimator, name, getattr(self, name))


def _check_constant_baselines(model):
    """Only ``ConstantBaseline`` is served by the device (K = mu psi)."""
    vec = getattr(model, '_baselines_vec', None)
    if vec is None:
        return     # the mirror: constant by construction
    for baseline in vec:
        basis = getattr(baseline, '_basis_mus', None)
        names = [type(b).__name__ for b in basis] if basis is not None else [type(baseline).__name__]
        if names != ['ConstantBaseline']:
            raise NotImplementedError('aslsd_b200 serves marked models with ConstantBaseline only '
                                      '(got %s)' % names)


class MTLHExactEstimB200(_B200Estimator):
    """Exact LSE_k and gradient of a marked model (MTLHExactEstim) on the device."""

    def initialize(self, k, n_iter, model, data, n_exact_single=None,
                   n_samples_adaptive_single=None, nonadaptive_sample_size_single=None,
                   single_strfs=None, n_samples_adaptive_double=None,
                   nonadaptive_sample_size_double=None, double_strfs=None):
        self.k = k
        self.n_iter = n_iter
        _check_constant_baselines(model)
        self.initialize_model_data(model, data)
        if self.list_marks is None:
            raise ValueError('a marked model needs ProcessPath.list_marks')
        self.stratif = {}
        self.list_log_names = self.get_list_log_names()
        self.make_param_indexing()
        self.initialize_logs()

    # ---- naming / indexing (mtlh_exact_estimator.py:181-251) ---------------------------------
    def get_list_log_names(self):
        d = self.d
        names = ['mu']
        for i in range(d):
            names += ['phi_%d' % i, 'upsilonzero_%d' % i, 'K_%d' % i]
        names += ['upsilon_%d_%d' % (i, j) for i, j in itertools.product(range(d), range(d))]
        return names

    def get_singlesum_lognames(self):
        d = self.d
        return ['mu'] + ['K_%d' % i for i in range(d)] + ['upsilonzero_%d' % i for i in range(d)]

    def make_param_indexing(self):
        d, k = self.d, self.k
        n_mu = self.vector_n_param_mu[k]
        nk, ni = self.matrix_n_param_ker[k], self.matrix_n_param_imp[k]
        self.n_f, self.vec_n_func, self.list_n_param = {}, {}, {}
        self.n_f['mu'], self.vec_n_func['mu'], self.list_n_param['mu'] = 1, [None], [n_mu]
        for i in range(d):
            for name in ('phi_%d' % i, 'upsilonzero_%d' % i):
                self.n_f[name], self.vec_n_func[name] = 2, [None, None]
                self.list_n_param[name] = [nk[i], ni[i]]
            name = 'K_%d' % i
            self.n_f[name], self.vec_n_func[name] = 2, [2, None]
            self.list_n_param[name] = [[None, nk[i], n_mu], ni[i]]
        for i, j in itertools.product(range(d), range(d)):
            name = 'upsilon_%d_%d' % (i, j)
            self.n_f[name] = 2
            if i == j:
                self.vec_n_func[name] = [None, None]
                self.list_n_param[name] = [nk[i], ni[i]]
            else:
                self.vec_n_func[name] = [2, 2]
                self.list_n_param[name] = [[None, nk[i], nk[j]], [None, ni[i], ni[j]]]

    def make_null_sumarray(self, n_f, list_n_param, vec_n_func=None, n_strata=None):
        """mtlh_exact_estimator.py:303-319"""
        def zeros(n):
            return np.zeros(n) if n_strata is None else np.zeros((n, n_strata))
        out = [None] * n_f
        for ix_f in range(n_f):
            if vec_n_func[ix_f] is None:
                out[ix_f] = zeros(list_n_param[ix_f])
            else:
                out[ix_f] = [None, zeros(list_n_param[ix_f][1]), zeros(list_n_param[ix_f][2])]
        return out

    def initialize_logs(self):
        self.logged_ixs = {name: {'adaptive': None, 'nonadaptive': None} for name in self.list_log_names}
        self.logged_allocs = {name: None for name in self.list_log_names}
        if self.is_log_lse:
            self.logged_lse = np.zeros(self.n_iter + 1)
        if self.is_log_total_estimates:
            self.M_term_log = np.zeros(self.n_iter)
            self.sum_f_log = {name: np.zeros(self.n_iter) for name in self.list_log_names}

    def clear_estimates(self):
        k = self.k
        self.M_term = 0.
        self.diff_M_term = np.zeros(self.vector_n_param_mu[k])
        for name in self.list_log_names:
            self.sum_f[name] = 0.
            self.sum_diff_f[name] = self.make_null_sumarray(self.n_f[name], self.list_n_param[name],
                                                            vec_n_func=self.vec_n_func[name])

    def log_estimates(self):
        if self.is_log_total_estimates:
            self.M_term_log[self.t] = self.M_term
            for name in self.list_log_names:
                self.sum_f_log[name][self.t] = self.sum_f[name]

    # ---- the sums ----------------------------------------------------------------------------
    def compute_objective(self, x_k, compute_f_sum=False, compute_diff_f_sum=True, count_iter=True,
                          verbose=False):
        """mtlh_exact_estimator.py:973-1023: host part O(N) (baseline, impact values), device
        part all pair and per-event sums."""
        self.clear_estimates()
        d, k = self.d, self.k
        x = np.asarray(x_k, dtype=np.float64)
        x_mu, x_ker, x_imp = self.xk2matrix_params(k, x)
        x_mu = np.asarray(x_mu, dtype=float)
        n_imp = [len(x_imp[i]) for i in range(d)]
        # M term and mu sum (:674-705) with the model's own baseline closures
        if compute_f_sum:
            self.M_term = float(self.M[k](self.T_f, x_mu))
            self.sum_f['mu'] = float(np.sum(self.mu[k](self.list_times[k], x_mu)))
        if compute_diff_f_sum:
            for r in range(len(x_mu)):
                self.diff_M_term[r] = float(self.diff_M[k](self.T_f, r, x_mu))
                self.sum_diff_f['mu'][0][r] = float(np.sum(self.diff_mu[k](self.list_times[k], r, x_mu)))
        if not (compute_f_sum or compute_diff_f_sum):
            return
        # impact values and derivatives per event, by the model's own closures
        marks = [np.asarray(m, dtype=float) for m in self.list_marks]
        w = [np.asarray(self.impact[k][i](marks[i], np.asarray(x_imp[i], dtype=float)), dtype=float)
             * np.ones(len(marks[i])) for i in range(d)]
        dw = [[np.asarray(self.diff_impact[k][i](marks[i], r, np.asarray(x_imp[i], dtype=float)),
                          dtype=float) * np.ones(len(marks[i]))
               for r in range(n_imp[i])] for i in range(d)]
        vectors, ix_w, ix_w2, ix_dw, ix_2wdw = [], [], [], [], []
        for i in range(d):
            ix_w.append(len(vectors)); vectors.append((i, w[i]))
            ix_w2.append(len(vectors)); vectors.append((i, w[i] ** 2))
            ix_dw.append([]); ix_2wdw.append([])
            if compute_diff_f_sum:
                for r in range(n_imp[i]):
                    ix_dw[i].append(len(vectors)); vectors.append((i, dw[i][r]))
                    ix_2wdw[i].append(len(vectors)); vectors.append((i, 2. * w[i] * dw[i][r]))
        engine = self._bound_engine()
        engine.set_weights(vectors)
        n_dev = 1 + sum(self.matrix_n_param_ker[k])
        x_dev = x[:n_dev]
        layout, _ = term_layout(engine.desc, k)
        term_ix = {entry[0]: t for t, entry in enumerate(layout)}
        # role rows: phi sources, upsilon targets, upsilon sources, psi (K) targets, upsilonzero targets
        base = np.array([ix_w, ix_w, ix_w, ix_w, ix_w2], dtype=np.int32)
        packed = engine.exact_sums_weighted(k, x_dev, base, compute_f_sum=True,
                                            compute_diff_f_sum=compute_diff_f_sum)
        sf, sdf = unpack_terms(engine.desc, k, packed)
        for i in range(d):
            self.sum_f['phi_%d' % i] = sf['phi_%d' % i]
            self.sum_f['upsilonzero_%d' % i] = sf['upsilonzero_%d' % i]
            self.sum_f['K_%d' % i] = sf['psi_%d' % i]
            if compute_diff_f_sum:
                self.sum_diff_f['phi_%d' % i][0] = sdf['phi_%d' % i]
                self.sum_diff_f['upsilonzero_%d' % i][0] = sdf['upsilonzero_%d' % i]
                self.sum_diff_f['K_%d' % i][0][1] = sdf['psi_%d' % i][1]
                self.sum_diff_f['K_%d' % i][0][2] = sdf['psi_%d' % i][2]
        for i, j in itertools.product(range(d), range(d)):
            name = 'upsilon_%d_%d' % (i, j)
            self.sum_f[name] = sf[name]
            if compute_diff_f_sum:
                if i == j:
                    self.sum_diff_f[name][0] = sdf[name]
                else:
                    self.sum_diff_f[name][0][1] = sdf[name][1]
                    self.sum_diff_f[name][0][2] = sdf[name][2]
        # impact-parameter derivatives: value sums with one weight vector replaced
        if compute_diff_f_sum:
            for p in range(d):
                for r in range(n_imp[p]):
                    # (a) the impact of the *target* event of type p is differentiated
                    roles = base.copy()
                    roles[1][p] = ix_dw[p][r]
                    roles[3][p] = ix_dw[p][r]
                    roles[4][p] = ix_2wdw[p][r]
                    mask = np.zeros(len(layout), dtype=np.uint8)
                    names = ['psi_%d' % p, 'upsilonzero_%d' % p] + ['upsilon_%d_%d' % (p, j) for j in range(d)]
                    for name in names:
                        mask[term_ix[name]] = 1
                    va, _ = unpack_terms(engine.desc, k, engine.exact_sums_weighted(
                        k, x_dev, roles, term_mask=mask, compute_f_sum=True, compute_diff_f_sum=False))
                    self.sum_diff_f['K_%d' % p][1][r] = va['psi_%d' % p]
                    self.sum_diff_f['upsilonzero_%d' % p][1][r] = va['upsilonzero_%d' % p]
                    for j in range(d):
                        if j != p:
                            self.sum_diff_f['upsilon_%d_%d' % (p, j)][1][1][r] = va['upsilon_%d_%d' % (p, j)]
                    # (b) the impact of the *source* event of type p is differentiated
                    roles = base.copy()
                    roles[0][p] = ix_dw[p][r]
                    roles[2][p] = ix_dw[p][r]
                    mask = np.zeros(len(layout), dtype=np.uint8)
                    for name in ['phi_%d' % p] + ['upsilon_%d_%d' % (i, p) for i in range(d)]:
                        mask[term_ix[name]] = 1
                    vb, _ = unpack_terms(engine.desc, k, engine.exact_sums_weighted(
                        k, x_dev, roles, term_mask=mask, compute_f_sum=True, compute_diff_f_sum=False))
                    self.sum_diff_f['phi_%d' % p][1][r] = vb['phi_%d' % p]
                    for i in range(d):
                        if i != p:
                            self.sum_diff_f['upsilon_%d_%d' % (i, p)][1][2][r] = vb['upsilon_%d_%d' % (i, p)]
                    # same type in both slots: product rule (:963-971)
                    self.sum_diff_f['upsilon_%d_%d' % (p, p)][1][r] = \
                        va['upsilon_%d_%d' % (p, p)] + vb['upsilon_%d_%d' % (p, p)]
        self.log_estimates()
        if count_iter:
            self.t += 1

    # ---- recombination (mtlh_exact_estimator.py:509-670) --------------------------------------
    def recombine_lse_k(self):
        d, T_f = self.d, self.T_f
        res = self.M_term - 2. * (self.sum_f['mu'] / T_f)
        for i in range(d):
            res += 2. * (self.sum_f['K_%d' % i] / T_f)
        for i, j in itertools.product(range(d), range(d)):
            res += 2. * (self.sum_f['upsilon_%d_%d' % (i, j)] / T_f)
        for j in range(d):
            res -= 2. * (self.sum_f['phi_%d' % j] / T_f)
        for i in range(d):
            res += self.sum_f['upsilonzero_%d' % i] / T_f
        return res

    def recombine_diff_lse_mu_kr(self, r):
        res = self.diff_M_term[r] - 2. * (self.sum_diff_f['mu'][0][r] / self.T_f)
        for i in range(self.d):
            res += 2. * (self.sum_diff_f['K_%d' % i][0][2][r] / self.T_f)
        return res

    def _recombine_theta(self, p, r, ix_f):
        d, T_f = self.d, self.T_f
        res = 0.
        for i in itertools.chain(range(p), range(p + 1, d)):
            res += 2. * (self.sum_diff_f['upsilon_%d_%d' % (i, p)][ix_f][2][r] / T_f)
        for j in itertools.chain(range(p), range(p + 1, d)):
            res += 2. * (self.sum_diff_f['upsilon_%d_%d' % (p, j)][ix_f][1][r] / T_f)
        res += 2. * (self.sum_diff_f['upsilon_%d_%d' % (p, p)][ix_f][r] / T_f)
        res -= 2. * (self.sum_diff_f['phi_%d' % p][ix_f][r] / T_f)
        if ix_f == 0:
            res += 2. * (self.sum_diff_f['K_%d' % p][0][1][r] / T_f)
        else:
            res += 2. * (self.sum_diff_f['K_%d' % p][1][r] / T_f)
        res += self.sum_diff_f['upsilonzero_%d' % p][ix_f][r] / T_f
        return res

    def recombine_diff_lse_theta_kpr_kernel(self, p, r):
        return self._recombine_theta(p, r, 0)

    def recombine_diff_lse_theta_kpr_impact(self, p, r):
        return self._recombine_theta(p, r, 1)

    def recombine_grad_lse_k(self):
        """mtlh_exact_estimator.py:647-670"""
        d, k = self.d, self.k
        grad = np.zeros(self.n_param_k)
        for ix in range(self.vector_n_param_mu[k]):
            grad[ix] = self.recombine_diff_lse_mu_kr(ix)
        for p in range(d):
            if self.matrix_n_param_ker[k][p]:
                for ix in range(self.interval_map_ker[k][p][0], self.interval_map_ker[k][p][1]):
                    grad[ix] = self.recombine_diff_lse_theta_kpr_kernel(p, self.ix_map_ker[k][ix]['par'])
        for p in range(d):
            if self.matrix_n_param_imp[k][p]:
                for ix in range(self.interval_map_imp[k][p][0], self.interval_map_imp[k][p][1]):
                    grad[ix] = self.recombine_diff_lse_theta_kpr_impact(p, self.ix_map_imp[k][ix]['par'])
        return grad

    # ---- LSE and gradient (:1028-1053) ----------------------------------------------------------
    def lse_k_estimate(self, x_k, compute_diff_f_sum=False, count_iter=True, verbose=False):
        self.compute_objective(x_k, compute_f_sum=True, compute_diff_f_sum=compute_diff_f_sum,
                               count_iter=count_iter, verbose=verbose)
        self.lse_k = self.recombine_lse_k()
        return self.lse_k

    def lse_k_grad_estimate(self, x_k, grad_alloc=True, rng=None, seed=None, count_iter=True,
                            verbose=False):
        self.compute_objective(x_k, compute_f_sum=self.is_log_lse, compute_diff_f_sum=True,
                               verbose=verbose)
        grad = self.recombine_grad_lse_k()
        self.grad_lse_k = grad
        if self.is_log_lse:
            self.lse_k = self.recombine_lse_k()
            self.logged_lse[self.t] = self.lse_k
        return grad
