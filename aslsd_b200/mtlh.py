"""Marked, time-dependent-baseline linear Hawkes models (MTLH) on the B200: exact and stratified LSE sums.

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
  * baselines: ``ConstantBaseline`` (K = mu psi, M = mu^2: basis_baseline_constant.py:30-65) and, in
    the exact estimator, ``PieceConstBaseline`` (constant on time intervals,
    basis_baseline_piececonst.py): its K term is a combination of weighted psi sums up to the
    interval ends, evaluated on the device with the horizon moved there (``aslsd_set_horizon``);
    M, mu and their derivatives are the model's own closures (O(1) / O(N) numpy).  These two are
    the baselines the reference's own MTLHExactEstim can evaluate (LinearBaseline is abstract,
    CosineBaseline.make_K is empty: scripts/probe_reference_scope.py).
``MTLHStratEstimB200`` (below) is the stratified counterpart (MTLHStratEstim).

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
        # baselines: {'type': 'constant'} (default) or {'type': 'piececonst', 'beta': breakpoints}
        self.baselines = [_normalise_baseline(b) for b in model.get('baselines', [{'type': 'constant'}] * d)]
        self.vector_n_param_mu = [1 if b['type'] == 'constant' else len(b['beta']) - 1 for b in self.baselines]
        self.matrix_n_param_ker = self.desc.matrix_n_param
        self.matrix_n_param_imp = [[_impact_n_param(self.impacts[k][i]) for i in range(d)]
                                   for k in range(d)]
        self.interval_map_ker = [[None] * d for _ in range(d)]
        self.interval_map_imp = [[None] * d for _ in range(d)]
        self.ix_map_ker = [dict() for _ in range(d)]
        self.ix_map_imp = [dict() for _ in range(d)]
        self.n_param_k = [0] * d
        for k in range(d):
            pos = self.vector_n_param_mu[k]
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
        closures = [_baseline_closures(b) for b in self.baselines]
        self.mu = [c[0] for c in closures]
        self.diff_mu = [c[1] for c in closures]
        self.M = [c[2] for c in closures]
        self.diff_M = [c[3] for c in closures]
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


def _normalise_baseline(spec):
    """Plain description of a baseline; the breakpoints as PieceConstBaseline.process_beta keeps them
    (basis_baseline_piececonst.py:59-67: sorted, unique, non-negative, 0 first, +inf last)."""
    if spec['type'] == 'constant':
        return {'type': 'constant'}
    if spec['type'] != 'piececonst':
        raise NotImplementedError('baseline %r is outside the engine scope' % (spec['type'],))
    q = np.array(sorted(set(float(v) for v in spec['beta'])))
    q = q[q >= 0.]
    if len(q) == 0 or q[0] != 0.:
        q = np.insert(q, 0, 0.)
    if q[-1] != np.inf:
        q = np.append(q, np.inf)
    return {'type': 'piececonst', 'beta': q}


def _baseline_closures(spec):
    """(mu, diff_mu, M, diff_M) of the mirror model: ConstantBaseline (basis_baseline_constant.py:30-49)
    or PieceConstBaseline (basis_baseline_piececonst.py:108-170; M is the time average of mu^2)."""
    if spec['type'] == 'constant':
        return (lambda t, x_mu: x_mu[0] * np.ones(len(t)), lambda t, ix, x_mu: np.ones(len(t)),
                lambda t, x_mu: x_mu[0] ** 2, lambda t, ix, x_mu: 2. * x_mu[0])
    beta = spec['beta']
    sizes = beta[1:] - beta[:-1]

    def g(t):
        return np.digitize(t, beta) - 1

    def mu(t, x_mu):
        return np.asarray(x_mu, dtype=float)[g(t)]

    def diff_mu(t, ix, x_mu):
        t = np.asarray(t, dtype=float)
        return ((t >= beta[ix]) & (t <= beta[ix + 1])).astype(float)

    def big_m(t, x_mu):
        x_mu = np.asarray(x_mu, dtype=float)
        if t == 0.:
            return 0.
        g_t = int(g(t))
        res = x_mu[g_t] ** 2 * (t - beta[g_t])
        if g_t > 0:
            res += np.cumsum(x_mu ** 2 * sizes)[g_t - 1]
        return res / t

    def diff_m(t, ix, x_mu):
        if t == 0.:
            return 0.
        g_t = int(g(t))
        if g_t > ix:
            return 2. * x_mu[ix] * (beta[ix + 1] - beta[ix]) / t
        if g_t == ix:
            return 2. * x_mu[ix] * (t - beta[ix]) / t
        return 0.
    return mu, diff_mu, big_m, diff_m


def _baseline_spec(model, k, allow_piececonst):
    """The baseline of sub-problem k as plain data: {'type': 'constant'} or {'type': 'piececonst',
    'beta': breakpoints}.  Anything else - several basis baselines, fixed variables, baselines the
    reference cannot evaluate either - is refused."""
    vec = getattr(model, '_baselines_vec', None)
    if vec is None:
        spec = getattr(model, 'baselines', None)
        spec = spec[k] if spec is not None else {'type': 'constant'}     # the mirror
    else:
        baseline = vec[k]
        basis = getattr(baseline, '_basis_mus', None)
        basis = list(basis) if basis is not None else [baseline]
        names = [type(b).__name__ for b in basis]
        if names == ['ConstantBaseline']:
            spec = {'type': 'constant'}
        elif names == ['PieceConstBaseline'] and not getattr(basis[0], 'n_fixed_vars', 0):
            spec = {'type': 'piececonst', 'beta': np.asarray(basis[0].beta, dtype=float)}
        else:
            raise NotImplementedError('aslsd_b200 serves marked models with one ConstantBaseline or one '
                                      'PieceConstBaseline without fixed variables (got %s)' % names)
    if spec['type'] != 'constant' and not allow_piececonst:
        raise NotImplementedError('the stratified marked estimator serves ConstantBaseline only '
                                  '(use MTLHExactEstimB200 for %s)' % spec['type'])
    return spec


class MTLHExactEstimB200(_B200Estimator):
    """Exact LSE_k and gradient of a marked model (MTLHExactEstim) on the device."""

    # save / load: the file layout of mtlh_exact_estimator.py:1058-1209 with the attributes this
    # estimator really has (upstream lists mu_term / diff_mu_term, which its marked estimators
    # never set, so its own save raises AttributeError)
    _STATE_ATTRS = ('t', 'M_term', 'diff_M_term', 'sum_f', 'sum_diff_f')
    _TOTAL_LOG_ATTRS = ('M_term_log', 'sum_f_log')
    _STRATIFICATIONS_OPTIONAL = True
    _ALLOW_PIECECONST = True

    def initialize_model_data(self, model, data):
        # (initialize and load both come through here)
        self.baseline_spec = _baseline_spec(model, self.k, allow_piececonst=self._ALLOW_PIECECONST)
        super().initialize_model_data(model, data)

    def _compute_terms(self, *args, **kwargs):
        raise NotImplementedError('marked models: the terms are evaluated together by compute_objective '
                                  '(one device call); the per-term entry points are not served')

    def initialize(self, k, n_iter, model, data, n_exact_single=None,
                   n_samples_adaptive_single=None, nonadaptive_sample_size_single=None,
                   single_strfs=None, n_samples_adaptive_double=None,
                   nonadaptive_sample_size_double=None, double_strfs=None):
        self.k = k
        self.n_iter = n_iter
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
        # the device's parameter vector: [mu, kernel parameters]; a piecewise-constant baseline has
        # several mu's - its K term is assembled below from sums taken with mu = 1
        n_mu = self.vector_n_param_mu[k]
        piececonst = self.baseline_spec['type'] == 'piececonst'
        x_dev = np.concatenate([[1. if piececonst else x[0]], x[n_mu:n_mu + sum(self.matrix_n_param_ker[k])]])
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
        if piececonst:
            self._piececonst_K(engine, x_dev, x_mu, w, dw, compute_diff_f_sum)
        self.log_estimates()
        if count_iter:
            self.t += 1

    def _piececonst_K(self, engine, x_dev, x_mu, w, dw, compute_diff_f_sum):
        """K_i sums of a baseline that is constant on the intervals [beta_j, beta_j+1)
        (PieceConstBaseline.make_K / make_diff_K, basis_baseline_piececonst.py:172-300, evaluated at
        t = T_f - t_n, s = t_n as compute_K_sum does, mtlh_exact_estimator.py:707-753).

        With g_s the interval of the event, g_ts that of t + s (the last one, G) and psi the integrated
        kernel, the reference's per-event value is
            g_s = G:  mu_G psi(t)
            g_s < G:  mu_gs psi(beta_gs+1 - s) + mu_G (psi(t) - psi(beta_G))            [sic: no - s]
                      + sum_{g_s < j < G} mu_j (psi(beta_j+1 - s) - psi(beta_j - s))
        so that, with Q_j = sum over the events before beta_j+1 of weight x psi(beta_j+1 - t_n) (j < G)
        and Q_G the same over all events up to T_f,
            K = sum_{j < G} mu_j (Q_j - Q_j-1) + mu_G (Q_G - psi(beta_G) x weight of the events before beta_G).
        Every Q_j is the device's weighted psi sum with the horizon moved to beta_j+1 and the targets
        restricted to the events before it; psi(beta_G) and its derivatives are the same sum over one
        unweighted event.  Derivatives: kernel parameters through d psi, mu_j through the coefficients,
        impact parameters through the weights."""
        d, k, T_f = self.d, self.k, self.T_f
        beta = np.asarray(self.baseline_spec['beta'], dtype=float)
        layout, _ = term_layout(engine.desc, k)
        term_ix = {entry[0]: t for t, entry in enumerate(layout)}
        for i in range(d):
            name = 'K_%d' % i
            times = np.asarray(self.list_times[i], dtype=float)
            n = len(times)
            nk, ni = self.matrix_n_param_ker[k][i], len(dw[i])
            val = 0.
            d_ker, d_mu, d_imp = np.zeros(nk), np.zeros(len(x_mu)), np.zeros(ni)
            if n == 0:
                continue
            mask = np.zeros(len(layout), dtype=np.uint8)
            mask[term_ix['psi_%d' % i]] = 1
            g_s = np.digitize(times, beta) - 1
            g_ts = np.digitize((T_f - times) + times, beta) - 1      # as the reference forms t + s

            def psi_sum(horizon, hi, vec, diff):
                """(sum, kernel derivatives) of weight x psi(horizon - t_n) over the first hi events."""
                if hi <= 0:
                    return 0., np.zeros(nk)
                engine.set_horizon(horizon)
                roles = -np.ones((5, d), dtype=np.int32)
                roles[3][i] = vec
                ranges = [(0, 0)] * d
                ranges[i] = (0, int(hi))
                sf, sdf = unpack_terms(engine.desc, k, engine.exact_sums_weighted(
                    k, x_dev, roles, term_mask=mask, compute_f_sum=True, compute_diff_f_sum=diff,
                    ranges=ranges))
                return sf['psi_%d' % i], (np.asarray(sdf['psi_%d' % i][1], dtype=float) if diff else np.zeros(nk))

            try:
                for G in np.unique(g_ts):
                    G = int(G)
                    ind = (g_ts == G).astype(float)
                    vectors = [(i, w[i] * ind)] + [(i, dw[i][r] * ind) for r in range(ni)]
                    engine.set_weights(vectors)
                    before_G = g_s < G
                    # psi(beta_G) and its derivatives: one unweighted event at distance beta_G of the horizon
                    psi_G, dpsi_G = (0., np.zeros(nk))
                    if G > 0 and np.any(before_G):
                        psi_G, dpsi_G = psi_sum(times[0] + beta[G], 1, -1, compute_diff_f_sum)
                    for v in range(1 + (ni if compute_diff_f_sum else 0)):
                        diff = compute_diff_f_sum and v == 0
                        weight = vectors[v][1]
                        prev, dprev = 0., np.zeros(nk)
                        for j in range(G + 1):
                            if j < G:
                                hi = int(np.searchsorted(times, beta[j + 1], side='left'))
                                q, dq = psi_sum(beta[j + 1], hi, v, diff)
                                c, dc = q - prev, dq - dprev
                                prev, dprev = q, dq
                            else:
                                q, dq = psi_sum(T_f, n, v, diff)
                                w_before = float(np.sum(weight[before_G]))
                                c, dc = q - psi_G * w_before, dq - dpsi_G * w_before
                            if v == 0:
                                val += x_mu[j] * c
                                d_ker += x_mu[j] * dc
                                d_mu[j] += c
                            else:
                                d_imp[v - 1] += x_mu[j] * c
            finally:
                engine.set_horizon(T_f)
            self.sum_f[name] = val
            if compute_diff_f_sum:
                self.sum_diff_f[name][0][1] = d_ker
                self.sum_diff_f[name][0][2] = d_mu
                self.sum_diff_f[name][1] = d_imp

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


class MTLHStratEstimB200(MTLHExactEstimB200):
    """Stratified (ASLSD) estimates of LSE_k and its gradient for a marked model
    (MTLHStratEstim, optimize/estimators/mtlh_stratified_estimator.py).

    Same division of labour as the exact estimator: index sampling (the reference's RNG call
    sequence, ``aslsd_b200.sampling``), impact values per event and the baseline term on the
    host; every sampled kernel functional, its product with the impact functional, the
    per-stratum sums, the allocation statistic (np.std of f per adaptive stratum, :719-724) and
    the allocation update on the device (``aslsd_sampled_sums_weighted``).  The derivative
    blocks with respect to the impact parameters are value-only launches on the same samples
    with one weight vector replaced by the derivative of the impact.

    Scope: what runs in the reference - ``grad_alloc=False`` (the ``fit`` default,
    mtlh.py:2428-2429; ``True`` fails upstream with an object-dtype cast, :720) and one
    adaptive iteration per call (``n_adaptive_iters == 1``, the stratifications' default; the
    multi-iteration branch indexes an integer, :770).
    """

    KIND = {'K': 0, 'upsilonzero': 1, 'phi': 2, 'upsilon': 3}
    _STATE_ATTRS = MTLHExactEstimB200._STATE_ATTRS + (
        'vecsum_f_adaptive', 'sum_f_adaptive', 'vecsum_f_nonadaptive', 'sum_f_nonadaptive', 'sum_f_exact',
        'vecsum_diff_f_adaptive', 'sum_diff_f_adaptive', 'vecsum_diff_f_nonadaptive',
        'sum_diff_f_nonadaptive', 'sum_diff_f_exact')
    _STRATIFICATIONS_OPTIONAL = False
    _ALLOW_PIECECONST = False

    # ---- stratifications (:76-169) -----------------------------------------------------------
    def initialize(self, k, n_iter, model, data, n_exact_single=None,
                   n_samples_adaptive_single=None, nonadaptive_sample_size_single=None,
                   single_strfs=None, n_samples_adaptive_double=None,
                   nonadaptive_sample_size_double=None, double_strfs=None):
        from aslsd_b200.stratifications import DoubleIxStratification, GeneralStratification
        self.k = k
        self.n_iter = n_iter
        self.initialize_model_data(model, data)
        if self.list_marks is None:
            raise ValueError('a marked model needs ProcessPath.list_marks')
        d = self.d
        single = {}
        for key, val in (('n_exact', n_exact_single), ('n_samples_adaptive', n_samples_adaptive_single),
                         ('nonadaptive_sample_size', nonadaptive_sample_size_single)):
            if val is not None:
                single[key] = val
        double = {}
        for key, val in (('n_samples_adaptive', n_samples_adaptive_double),
                         ('nonadaptive_sample_size', nonadaptive_sample_size_double)):
            if val is not None:
                double[key] = val
        single_strfs = single_strfs or {}
        double_strfs = double_strfs or {}
        self.stratif = {}
        self.stratif['mu'] = single_strfs.get('mu_stratification') or \
            GeneralStratification(n=self.n_events[k], **single)
        for key, prefix in (('K_stratification', 'K_'), ('upsilonzero_stratification', 'upsilonzero_')):
            given = single_strfs.get(key, [None] * d)
            for i in range(d):
                self.stratif[prefix + str(i)] = given[i] or GeneralStratification(n=self.n_events[i],
                                                                                  **single)
        given = double_strfs.get('upsilon_stratification', [[None] * d for _ in range(d)])
        for i, j in itertools.product(range(d), range(d)):
            strf = given[i][j]
            if strf is None:
                strf = DoubleIxStratification(h_max=self.kappa[j][i][self.n_events[i] - 1] + 1, **double)
                strf.get_stratification_size(self.lag_sizes, i, j)
            self.stratif['upsilon_%d_%d' % (i, j)] = strf
        given = double_strfs.get('phi_stratification', [None] * d)
        for i in range(d):
            strf = given[i]
            if strf is None:
                strf = DoubleIxStratification(h_max=self.kappa[i][k][self.n_events[k] - 1] + 1, **double)
                strf.get_stratification_size(self.lag_sizes, k, i)
            self.stratif['phi_%d' % i] = strf
        for name, strf in self.stratif.items():
            if getattr(strf, 'n_adaptive_iters', 1) != 1:
                raise NotImplementedError('stratification %s: several adaptive iterations per call '
                                          'are not served (they fail upstream too)' % name)
        self.list_log_names = self.get_list_log_names()
        self.make_param_indexing()
        self.initialize_logs()
        if self.is_log_ixs:
            self.logged_ixs = {name: {'adaptive': [None] * n_iter, 'nonadaptive': [None] * n_iter}
                               for name in self.list_log_names}
        if self.is_log_allocs:
            self.logged_allocs = {name: [None] * n_iter for name in self.list_log_names}

    def term_list(self):
        """(log_name, family, i, j) in the order compute_objective evaluates - and draws
        samples for - the terms (:1648-1700)."""
        d = self.d
        terms = [('mu', 'mu', self.k, self.k)]
        for i in range(d):
            terms += [('K_%d' % i, 'K', i, i), ('upsilonzero_%d' % i, 'upsilonzero', i, i),
                      ('phi_%d' % i, 'phi', i, i), ('upsilon_%d_%d' % (i, i), 'upsilon', i, i)]
        for i, j in itertools.product(range(d), range(d)):
            if i != j:
                terms.append(('upsilon_%d_%d' % (i, j), 'upsilon', i, j))
        return terms

    def clear_estimates(self):
        super().clear_estimates()
        null = self.make_null_sumarray
        for attr in ('vecsum_f_adaptive', 'sum_f_adaptive', 'vecsum_f_nonadaptive', 'sum_f_nonadaptive',
                     'sum_f_exact', 'vecsum_diff_f_adaptive', 'sum_diff_f_adaptive',
                     'vecsum_diff_f_nonadaptive', 'sum_diff_f_nonadaptive', 'sum_diff_f_exact'):
            setattr(self, attr, {})
        for name in self.list_log_names:
            strf = self.stratif[name]
            args = (self.n_f[name], self.list_n_param[name])
            kw = dict(vec_n_func=self.vec_n_func[name])
            self.vecsum_f_adaptive[name] = np.zeros(strf.n_adaptive_strata)
            self.vecsum_f_nonadaptive[name] = np.zeros(strf.n_nonadaptive_strata)
            self.sum_f_adaptive[name] = self.sum_f_nonadaptive[name] = self.sum_f_exact[name] = 0.
            self.vecsum_diff_f_adaptive[name] = null(*args, n_strata=strf.n_adaptive_strata, **kw)
            self.vecsum_diff_f_nonadaptive[name] = null(*args, n_strata=strf.n_nonadaptive_strata, **kw)
            self.sum_diff_f_adaptive[name] = null(*args, **kw)
            self.sum_diff_f_nonadaptive[name] = null(*args, **kw)
            self.sum_diff_f_exact[name] = null(*args, **kw)

    def _draw(self, family, i, j, strata, n_samples, rng):
        from aslsd_b200 import sampling
        if family in ('mu', 'K', 'upsilonzero'):
            raw = sampling.sample_single_ix_stratified(strata, n_samples, rng)
            return raw, [(np.asarray(ix), None) for ix in raw]
        tgt, src = (self.k, i) if family == 'phi' else (i, j)
        raw = sampling.sample_double_ix_stratified(tgt, src, strata, n_samples, self.kappa, self.varpi,
                                                   self.n_events, rng)
        return raw, [(np.asarray(s['m']), np.asarray(s['n'])) for s in raw]

    # ---- the estimates -------------------------------------------------------------------------
    def compute_objective(self, x_k, compute_f_sum=False, compute_diff_f_sum=True, grad_alloc=True,
                          rng=None, seed=None, count_iter=True):
        if grad_alloc:
            raise NotImplementedError('grad_alloc=True fails in the reference for marked models '
                                      '(mtlh_stratified_estimator.py:720); use grad_alloc=False')
        self.clear_estimates()
        if rng is None:
            rng = np.random.default_rng(seed)
        d, k = self.d, self.k
        x = np.asarray(x_k, dtype=np.float64)
        x_mu, x_ker, x_imp = self.xk2matrix_params(k, x)
        x_mu = np.asarray(x_mu, dtype=float)
        n_imp = [len(x_imp[i]) for i in range(d)]
        t_iter = self.t
        if compute_f_sum:
            self.M_term = float(self.M[k](self.T_f, x_mu))
        if compute_diff_f_sum:
            for r in range(len(x_mu)):
                self.diff_M_term[r] = float(self.diff_M[k](self.T_f, r, x_mu))
        # ---- 1. draw every sample, term after term (host RNG, reference order) ----------------
        plans = []
        for name, family, i, j in self.term_list():
            strf = self.stratif[name]
            plan = {'name': name, 'family': family, 'i': i, 'j': j, 'blocks': [], 'chunks': []}
            if strf.n_adaptive_strata > 0:
                n_a = strf.n_adaptive_strata
                rel_alloc = (strf.ema_weight * strf.adaptive_rel_alloc
                             + (1. - strf.ema_weight) * strf.guess_rel_alloc)           # :645-646
                strf.guess_rel_alloc = copy.deepcopy(rel_alloc)                         # :649
                n_samples = [min(strf.adaptive_strata_sizes[s],
                                 max(1, int(rel_alloc[s] * (n_a + strf.n_samples_adaptive[0]))))
                             for s in range(n_a)]                                       # :668
                raw, flat = self._draw(family, i, j, strf.adaptive_strata, n_samples, rng)
                if self.is_log_ixs:
                    self.logged_ixs[name]['adaptive'][t_iter] = [[s] for s in raw]
                plan['chunks'] += flat
                plan['blocks'].append(('adaptive', n_a))
                plan['default_alloc'] = np.asarray(rel_alloc, dtype=float)
            if strf.n_nonadaptive_strata > 0:
                raw, flat = self._draw(family, i, j, strf.nonadaptive_strata, strf.nonadaptive_abs_alloc, rng)
                if self.is_log_ixs:
                    self.logged_ixs[name]['nonadaptive'][t_iter] = raw
                plan['chunks'] += flat
                plan['blocks'].append(('nonadaptive', strf.n_nonadaptive_strata))
            if strf.n_exact > 0:
                plan['chunks'].append((np.asarray(strf.exact_domain), None))
                plan['blocks'].append(('exact', 1))
            plans.append(plan)
        # which sums every domain produces (:629-634, :893-894, :983-984): f is always needed
        # here (grad_alloc is False; n_func = 2 forces it with the derivatives)
        do_f = True
        do_diff = bool(compute_diff_f_sum)
        # ---- 2. baseline term on the host (constant in t: O(samples)) ---------------------------
        self._estimate_mu(plans[0], x_mu, compute_f_sum, do_diff)
        # ---- 3. impact values per event, weight vectors ------------------------------------------
        marks = [np.asarray(m, dtype=float) for m in self.list_marks]
        w = [np.asarray(self.impact[k][i](marks[i], np.asarray(x_imp[i], dtype=float)), dtype=float)
             * np.ones(len(marks[i])) for i in range(d)]
        dw = [[np.asarray(self.diff_impact[k][i](marks[i], r, np.asarray(x_imp[i], dtype=float)),
                          dtype=float) * np.ones(len(marks[i])) for r in range(n_imp[i])] for i in range(d)]
        vectors, ix_w, ix_w2, ix_dw, ix_2wdw = [], [], [], [], []
        for i in range(d):
            ix_w.append(len(vectors)); vectors.append((i, w[i]))
            ix_w2.append(len(vectors)); vectors.append((i, w[i] ** 2))
            ix_dw.append([]); ix_2wdw.append([])
            if do_diff:
                for r in range(n_imp[i]):
                    ix_dw[i].append(len(vectors)); vectors.append((i, dw[i][r]))
                    ix_2wdw[i].append(len(vectors)); vectors.append((i, 2. * w[i] * dw[i][r]))
        engine = self._bound_engine()
        engine.set_weights(vectors)
        x_dev = x[:1 + sum(self.matrix_n_param_ker[k])]
        base = np.array([ix_w, ix_w, ix_w, ix_w, ix_w2], dtype=np.int32)
        # ---- 4. one launch for every kernel-side output of every term ----------------------------
        dev_plans = [p for p in plans[1:] if p['chunks']]
        self._plans_cache = dev_plans
        res = self._launch(engine, x_dev, dev_plans, base, do_diff, with_alloc=True)
        for plan in dev_plans:
            self._store_kernel_side(plan, res[plan['name']], do_f, do_diff, t_iter)
        # ---- 5. impact-side derivatives: value-only launches with one weight replaced ------------
        if do_diff:
            by_name = {p['name']: p for p in dev_plans}
            for p_dim in range(d):
                for r in range(n_imp[p_dim]):
                    roles = base.copy()
                    roles[1][p_dim] = ix_dw[p_dim][r]
                    roles[3][p_dim] = ix_dw[p_dim][r]
                    roles[4][p_dim] = ix_2wdw[p_dim][r]
                    names = ['K_%d' % p_dim, 'upsilonzero_%d' % p_dim] + ['upsilon_%d_%d' % (p_dim, j) for j in range(d)]
                    sel = [by_name[n] for n in names if n in by_name]
                    va = self._launch(engine, x_dev, sel, roles, False, with_alloc=False)
                    roles = base.copy()
                    roles[0][p_dim] = ix_dw[p_dim][r]
                    roles[2][p_dim] = ix_dw[p_dim][r]
                    names = ['phi_%d' % p_dim] + ['upsilon_%d_%d' % (i, p_dim) for i in range(d)]
                    sel = [by_name[n] for n in names if n in by_name]
                    vb = self._launch(engine, x_dev, sel, roles, False, with_alloc=False)
                    self._store_impact_side(p_dim, r, va, vb)
        # ---- 6. totals (:1065-1094) ----------------------------------------------------------------
        for plan in plans[1:]:
            name = plan['name']
            if compute_f_sum:
                self.sum_f[name] = (self.sum_f_adaptive[name] + self.sum_f_nonadaptive[name]
                                    + self.sum_f_exact[name])
            if do_diff:
                self._add_domains(name)
        self.log_estimates()
        if count_iter:
            self.t += 1

    # per-stratum estimate: size * mean for sampled strata, plain sum on the exact domain
    def _scale(self, plan, block_sums):
        """block_sums: (n_strata_total, n_out) plain per-stratum sums in block order ->
        list of (domain, estimates (n_strata, n_out))."""
        strf = self.stratif[plan['name']]
        counts = np.array([len(m) for m, _ in plan['chunks']], dtype=float)
        out, s = [], 0
        for domain, n_strata in plan['blocks']:
            blk = block_sums[s:s + n_strata]
            if domain == 'exact':
                est = blk
            else:
                sizes = np.asarray(strf.adaptive_strata_sizes if domain == 'adaptive'
                                   else strf.nonadaptive_strata_sizes, dtype=float)
                est = (blk / counts[s:s + n_strata, None]) * sizes[:, None]
            out.append((domain, est))
            s += n_strata
        return out

    def _launch(self, engine, x_dev, sel_plans, roles, with_diff, with_alloc):
        """One sampled launch over the strata of ``sel_plans``; returns
        {name: (per-stratum sums (n_strata, n_out), stat_std, rel_alloc)}."""
        if not sel_plans:
            return {}
        dev_terms, offsets, all_m, all_n, sizes, defaults = [], [0], [], [], [], []
        for plan in sel_plans:
            strf = self.stratif[plan['name']]
            for m, n in plan['chunks']:
                all_m.append(m)
                all_n.append(n if n is not None else np.zeros(len(m), dtype=np.int64))
                offsets.append(offsets[-1] + len(m))
            for domain, n_strata in plan['blocks']:
                if domain == 'adaptive':
                    sizes += [float(s) for s in strf.adaptive_strata_sizes]
                    defaults += [float(a) for a in plan['default_alloc']]
                else:
                    sizes += [0.] * n_strata
                    defaults += [0.] * n_strata
            dev_terms.append((self.KIND[plan['family']], plan['i'], plan['j'], len(plan['chunks'])))
        sums, out_off, stat_mean, stat_std, rel = engine.sampled_sums(
            self.k, x_dev, dev_terms, offsets, np.concatenate(all_m), np.concatenate(all_n),
            strata_sizes=sizes if with_alloc else None, default_alloc=defaults if with_alloc else None,
            compute_diff_f_sum=with_diff, grad_alloc=False, roles=roles)
        out, s_ix = {}, 0
        for plan in sel_plans:
            n_s = len(plan['chunks'])
            block = np.stack([sums[out_off[s_ix + s]:out_off[s_ix + s + 1]] for s in range(n_s)])
            out[plan['name']] = (block, stat_std[s_ix:s_ix + n_s],
                                 None if rel is None else rel[s_ix:s_ix + n_s])
            s_ix += n_s
        return out

    def _store_kernel_side(self, plan, result, do_f, do_diff, t_iter):
        """Function value and kernel / baseline derivative blocks (ix_f = 0) of one term."""
        name, family = plan['name'], plan['family']
        strf = self.stratif[name]
        block, stat_std, rel = result
        nk_i = self.matrix_n_param_ker[self.k][plan['i']]
        s = 0
        for domain, est in self._scale(plan, block):
            n_strata = est.shape[0]
            if domain == 'exact':
                self.sum_f_exact[name] = est[0, 0]
            else:
                getattr(self, 'vecsum_f_' + domain)[name] = est[:, 0].copy()
                getattr(self, 'sum_f_' + domain)[name] = est[:, 0].sum()
            if do_diff:
                # device layout: [f, d/d kernel params of ix_func 1 (, ix_func 2 or baseline)]
                if family == 'K':
                    parts = {1: est[:, 1:1 + nk_i], 2: est[:, 1 + nk_i:2 + nk_i]}
                elif family == 'upsilon' and plan['i'] != plan['j']:
                    nk_j = self.matrix_n_param_ker[self.k][plan['j']]
                    parts = {1: est[:, 1:1 + nk_i], 2: est[:, 1 + nk_i:1 + nk_i + nk_j]}
                else:
                    parts = {None: est[:, 1:1 + nk_i]}
                for ix_func, arr in parts.items():
                    per_param = arr.T.copy()
                    self._put(name, domain, 0, ix_func, per_param)
            if domain == 'adaptive':
                strf.adaptive_rel_alloc = np.array(rel[s:s + n_strata], dtype=float)      # :861
                if self.is_log_allocs:
                    self.logged_allocs[name][t_iter] = strf.adaptive_rel_alloc.copy()
            s += n_strata

    def _put(self, name, domain, ix_f, ix_func, per_param):
        """per_param: (n_param, n_strata) estimates of one derivative block of one domain."""
        def assign(container, value):
            if ix_func is None:
                container[name][ix_f] = value
            else:
                container[name][ix_f][ix_func] = value
        if domain != 'exact':
            assign(getattr(self, 'vecsum_diff_f_' + domain), per_param)
        assign(getattr(self, 'sum_diff_f_' + domain), per_param.sum(axis=1))

    def _store_impact_side(self, p, r, va, vb):
        """Derivative with respect to the r-th impact parameter of I_kp (ix_f = 1) from the
        value-only launches va (target weight differentiated) and vb (source weight)."""
        d = self.d
        by_name = {}
        for res in (va, vb):
            for name, (block, _, _) in res.items():
                by_name.setdefault(name, []).append(block)

        def domains(name, blocks):
            plan = next(pl for pl in self._plans_cache if pl['name'] == name)
            total = sum(blocks)
            return self._scale(plan, total)
        for name, blocks in by_name.items():
            family = name.split('_')[0]
            for domain, est in domains(name, blocks):
                vals = est[:, 0]                 # (n_strata,) estimates of sum f x dI
                if family == 'upsilon':
                    i, j = (int(v) for v in name.split('_')[1:])
                    if i == j:
                        self._put_row(name, domain, 1, None, r, vals)
                    else:
                        self._put_row(name, domain, 1, 1 if i == p else 2, r, vals)
                else:
                    self._put_row(name, domain, 1, None, r, vals)

    def _put_row(self, name, domain, ix_f, ix_func, r, vals):
        def target(container):
            return container[name][ix_f] if ix_func is None else container[name][ix_f][ix_func]
        if domain != 'exact':
            target(getattr(self, 'vecsum_diff_f_' + domain))[r] = vals
        target(getattr(self, 'sum_diff_f_' + domain))[r] = vals.sum()

    def _add_domains(self, name):
        parts = (self.sum_diff_f_adaptive[name], self.sum_diff_f_nonadaptive[name], self.sum_diff_f_exact[name])
        for ix_f in range(self.n_f[name]):
            if self.vec_n_func[name][ix_f] is None:
                self.sum_diff_f[name][ix_f] = parts[0][ix_f] + parts[1][ix_f] + parts[2][ix_f]
            else:
                for ix_func in (1, 2):
                    self.sum_diff_f[name][ix_f][ix_func] = (parts[0][ix_f][ix_func] + parts[1][ix_f][ix_func]
                                                            + parts[2][ix_f][ix_func])

    def _estimate_mu(self, plan, x_mu, compute_f_sum, do_diff):
        """estimate_sum of the 'mu' term (:1273-1316) with the model's own baseline closures."""
        name = 'mu'
        strf = self.stratif[name]
        t_k = np.asarray(self.list_times[self.k])
        n_mu = len(x_mu)
        s = 0
        for domain, n_strata in plan['blocks']:
            f_est = np.zeros(n_strata)
            d_est = np.zeros((n_mu, n_strata))
            std = np.zeros(n_strata)
            for q in range(n_strata):
                ix = plan['chunks'][s + q][0]
                f_vals = np.asarray(self.mu[self.k](t_k[ix], x_mu), dtype=float) * np.ones(len(ix))
                scale = 1.
                if domain == 'adaptive':
                    scale = strf.adaptive_strata_sizes[q]
                elif domain == 'nonadaptive':
                    scale = strf.nonadaptive_strata_sizes[q]
                f_est[q] = np.sum(f_vals) if domain == 'exact' else scale * np.mean(f_vals)
                std[q] = np.std(f_vals)
                for r in range(n_mu):
                    dv = np.asarray(self.diff_mu[self.k](t_k[ix], r, x_mu), dtype=float) * np.ones(len(ix))
                    d_est[r, q] = np.sum(dv) if domain == 'exact' else scale * np.mean(dv)
            if domain == 'exact':
                self.sum_f_exact[name] = f_est[0]
            else:
                getattr(self, 'vecsum_f_' + domain)[name] = f_est
                getattr(self, 'sum_f_' + domain)[name] = f_est.sum()
            if do_diff:
                self._put(name, domain, 0, None, d_est)
            if domain == 'adaptive':
                sizes = np.asarray(strf.adaptive_strata_sizes, dtype=float)
                with np.errstate(divide='ignore', invalid='ignore'):
                    rel = sizes * np.sqrt(sizes / (sizes - 1.)) * std                     # :725-729
                tot = np.sum(rel)
                strf.adaptive_rel_alloc = rel / tot if tot > 0. else plan['default_alloc'].copy()
                if self.is_log_allocs:
                    self.logged_allocs[name][self.t] = strf.adaptive_rel_alloc.copy()
            s += n_strata
        if compute_f_sum:
            self.sum_f[name] = self.sum_f_adaptive[name] + self.sum_f_nonadaptive[name] + self.sum_f_exact[name]
        if do_diff:
            self._add_domains(name)

    # ---- LSE and gradient (:1708-1745) -------------------------------------------------------------
    def lse_k_estimate(self, x_k, compute_diff_f_sum=False, grad_alloc=True, rng=None, seed=None,
                       count_iter=True):
        if rng is None:
            rng = np.random.default_rng(seed)
        self.compute_objective(x_k, compute_f_sum=True, compute_diff_f_sum=compute_diff_f_sum,
                               grad_alloc=grad_alloc, rng=rng, count_iter=count_iter)
        self.lse_k = self.recombine_lse_k()
        return self.lse_k

    def lse_k_grad_estimate(self, x_k, grad_alloc=True, rng=None, seed=None, count_iter=True):
        if rng is None:
            rng = np.random.default_rng(seed)
        self.compute_objective(x_k, compute_f_sum=self.is_log_lse, compute_diff_f_sum=True,
                               grad_alloc=grad_alloc, rng=rng, count_iter=count_iter)
        grad = self.recombine_grad_lse_k()
        self.grad_lse_k = grad
        if self.is_log_lse:
            self.lse_k = self.recombine_lse_k()
            self.logged_lse[self.t] = self.lse_k
        return grad
