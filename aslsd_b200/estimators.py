"""LSE estimators of the reference's plug-in API, evaluated on the B200.

``MHPExactEstimB200`` and ``MHPStratEstimB200`` expose the interface the
reference's fit loop drives (``MHP.fit(..., estimators=[...])``,
models/hawkes/linear/mhp.py:1722-1787):

    est.initialize(k, n_iter, model, process_path, n_exact_single=..., ...)
    g = est.lse_k_grad_estimate(x_k, rng=rng, grad_alloc=bool)
    v = est.lse_k_estimate(x_k)
    est.get_log(), est.logged_lse, est.t

and leave the same public state behind (``sum_f`` / ``sum_diff_f`` with the
shapes of ``make_null_sumarray``, ``vecsum_*`` / ``sum_*_{adaptive,nonadaptive,
exact}``, ``stratif``, logs), so the reference's ``recombine_*`` arithmetic,
loggers and pickling keep working.  Interfaces mirrored:

  * Estimator ABC          optimize/estimators/estimator.py:6-57
  * MHPExactEstim          optimize/estimators/mhp_exact_estimator.py
        compute_objective :709-761, recombine :442-530, lse_k_* :765-788
  * MHPStratEstim          optimize/estimators/mhp_stratified_estimator.py
        set_stratification :55-204, estimate_sum_* :534-971,
        compute_objective :1333-1386, recombine :976-1068, lse_k_* :1391-1427

``model`` may be the reference's own ``MHP`` or the mirror of
``aslsd_b200.model``; ``process_path`` the reference's ``ProcessPath`` or
``aslsd_b200.ProcessPath`` (only their data attributes are read).  When the
reference package is importable these classes derive from its ``Estimator``
ABC, so a single instance can be passed to ``fit`` and is deep-copied per
dimension exactly like a reference estimator (mhp.py:1739-1741).

All sums are computed by the CUDA library (``aslsd_b200.engine.Engine``): one
call per ``compute_objective``.  There is no CPU fallback.
"""
import copy
import itertools
import os
import pickle

import numpy as np

from aslsd_b200 import sampling, saturation
from aslsd_b200.engine import unpack_terms
from aslsd_b200.stratifications import DoubleIxStratification, GeneralStratification

try:   # drop-in: be a subclass of the reference's ABC when it is installed
    from aslsd.optimize.estimators.estimator import Estimator as _RefEstimator
except Exception:   # the reference package is absent (e.g. on the GPU box)
    _RefEstimator = object

KIND_PSI, KIND_UPSILONZERO, KIND_PHI, KIND_UPSILON = 0, 1, 2, 3


def _engine_for(data):
    from aslsd_b200.events import engine_for
    return engine_for(data)


class _B200Estimator(_RefEstimator):
    """State and behaviour shared by the two estimators."""

    def __init__(self, is_grad_target=False, is_log_ixs=False, is_log_allocs=False,
                 is_log_total_estimates=False, is_log_strata_estimates=False,
                 is_log_lse=False, engine=None):
        self.t = 0
        self.is_grad_target = is_grad_target
        self.is_log_ixs = is_log_ixs
        self.is_log_allocs = is_log_allocs
        self.is_log_total_estimates = is_log_total_estimates
        self.is_log_strata_estimates = is_log_strata_estimates
        self.is_log_lse = is_log_lse
        self.engine = engine
        self.sum_f = {}
        self.sum_diff_f = {}

    # device handles are not picklable state (checkpoint row):
    # they are rebuilt on the next initialize_model_data / bind
    def __getstate__(self):
        state = dict(self.__dict__)
        state['engine'] = None
        state.pop('_model', None)
        state.pop('_path', None)
        return state

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for key, val in self.__dict__.items():
            # the d copies made by fit share the engine (one device context)
            new.__dict__[key] = val if key in ('engine', '_model', '_path') else copy.deepcopy(val, memo)
        return new

    def initialize_model_data(self, model, data):
        """estimator.py:30-45 plus the binding of the device context."""
        self.list_times = data.list_times
        self.list_marks = getattr(data, 'list_marks', None)
        self.T_f = data.T_f
        self.d = data.d
        self.n_events = data.n_events
        self.eta = data.eta
        self.list_times2end = data.list_times2end
        self.kappa = data.kappa
        self.varpi = data.varpi
        self.lag_sizes = data.lag_sizes
        model.init_estimator(self, self.k)
        if self.engine is None:
            self.engine = _engine_for(data)
        self._model, self._path = model, data
        self.desc = self.engine.bind(data, model)

    def _bound_engine(self):
        if self.engine is None:
            raise RuntimeError('estimator is not bound to a device context: call initialize '
                               '(the B200 engine has no CPU fallback)')
        self.engine.bind(self._path, self._model)
        return self.engine

    # ---- naming / indexing (mhp_exact_estimator.py:119-170) -----------------------------
    def get_list_log_names(self):
        d = self.d
        names = []
        for i in range(d):
            names += ['phi_%d' % i, 'upsilonzero_%d' % i, 'psi_%d' % i]
        names += ['upsilon_%d_%d' % (i, j) for i, j in itertools.product(range(d), range(d))]
        return names

    def get_singlesum_lognames(self):
        d = self.d
        return ['psi_%d' % i for i in range(d)] + ['upsilonzero_%d' % i for i in range(d)]

    def get_doublesum_lognames(self):
        d = self.d
        return (['upsilon_%d_%d' % (i, j) for i, j in itertools.product(range(d), range(d))]
                + ['phi_%d' % i for i in range(d)])

    def make_param_indexing(self):
        d, k = self.d, self.k
        self.n_func, self.n_param = {}, {}
        for i in range(d):
            n_i = self.matrix_n_param[k][i]
            self.n_func['phi_%d' % i], self.n_param['phi_%d' % i] = 1, n_i
            self.n_func['upsilonzero_%d' % i], self.n_param['upsilonzero_%d' % i] = 1, n_i
            self.n_func['psi_%d' % i], self.n_param['psi_%d' % i] = 2, [None, n_i, 1]
        for i, j in itertools.product(range(d), range(d)):
            name = 'upsilon_%d_%d' % (i, j)
            if i == j:
                self.n_func[name], self.n_param[name] = 1, self.matrix_n_param[k][i]
            else:
                self.n_func[name] = 2
                self.n_param[name] = [None, self.matrix_n_param[k][i], self.matrix_n_param[k][j]]

    def make_null_sumarray(self, n_param, n_func=1, n_strata=None):
        def zeros(n):
            return np.zeros(n) if n_strata is None else np.zeros((n, n_strata))
        if n_func == 1:
            return zeros(n_param)
        return [None, zeros(n_param[1]), zeros(n_param[2])]

    def term_list(self):
        """(log_name, kind, i, j) in the order the reference evaluates - and
        draws samples for - the terms (mhp_stratified_estimator.py:1348-1380)."""
        d = self.d
        terms = []
        for i in range(d):
            terms += [('psi_%d' % i, KIND_PSI, i, i), ('upsilonzero_%d' % i, KIND_UPSILONZERO, i, i),
                      ('phi_%d' % i, KIND_PHI, i, i), ('upsilon_%d_%d' % (i, i), KIND_UPSILON, i, i)]
        for i, j in itertools.product(range(d), range(d)):
            if i != j:
                terms.append(('upsilon_%d_%d' % (i, j), KIND_UPSILON, i, j))
        return terms

    # ---- mu term and recombination ---------------------------------------------------------
    def compute_mu_term(self, mu_k, compute_f_sum=False, compute_diff_f_sum=True):
        k = self.k
        if compute_f_sum:
            self.mu_term = mu_k ** 2 - 2. * self.eta[k] * mu_k
        if compute_diff_f_sum:
            self.diff_mu_term = 2. * (mu_k - self.eta[k])

    def recombine_lse_k(self):
        """mhp_exact_estimator.py:442-465 / mhp_stratified_estimator.py:976-999."""
        d, T_f = self.d, self.T_f
        res = 0.
        res += self.mu_term
        for i in range(d):
            res += 2. * (self.sum_f['psi_%d' % i] / T_f)
        for i, j in itertools.product(range(d), range(d)):
            res += 2. * (self.sum_f['upsilon_%d_%d' % (i, j)] / T_f)
        for j in range(d):
            res -= 2. * (self.sum_f['phi_%d' % j] / T_f)
        for i in range(d):
            res += self.sum_f['upsilonzero_%d' % i] / T_f
        return res

    def recombine_diff_lse_mu_k(self):
        res = self.diff_mu_term
        for i in range(self.d):
            res += 2. * (self.sum_diff_f['psi_%d' % i][2][0] / self.T_f)
        return res

    def recombine_diff_lse_theta_kpr(self, p, r):
        """Derivative with respect to the r-th parameter of phi_kp
        (mhp_stratified_estimator.py:1017-1053)."""
        d, T_f = self.d, self.T_f
        res = 0.
        for i in itertools.chain(range(p), range(p + 1, d)):
            res += 2. * (self.sum_diff_f['upsilon_%d_%d' % (i, p)][2][r] / T_f)
        for j in itertools.chain(range(p), range(p + 1, d)):
            res += 2. * (self.sum_diff_f['upsilon_%d_%d' % (p, j)][1][r] / T_f)
        res += 2. * (self.sum_diff_f['upsilon_%d_%d' % (p, p)][r] / T_f)
        res -= 2. * (self.sum_diff_f['phi_%d' % p][r] / T_f)
        res += 2. * (self.sum_diff_f['psi_%d' % p][1][r] / T_f)
        res += self.sum_diff_f['upsilonzero_%d' % p][r] / T_f
        return res

    def _recombine_theta_block(self, p):
        """recombine_diff_lse_theta_kpr(p, r) for every parameter r of phi_kp at once: the same terms
        in the same order, element by element (bit-identical to the scalar loop; one estimator call
        at d = 10 made 100 of those calls, 2 400 dictionary look-ups)."""
        d, T_f = self.d, self.T_f
        res = np.zeros(self.matrix_n_param[self.k][p])
        for i in itertools.chain(range(p), range(p + 1, d)):
            res += 2. * (np.asarray(self.sum_diff_f['upsilon_%d_%d' % (i, p)][2], dtype=np.float64) / T_f)
        for j in itertools.chain(range(p), range(p + 1, d)):
            res += 2. * (np.asarray(self.sum_diff_f['upsilon_%d_%d' % (p, j)][1], dtype=np.float64) / T_f)
        res += 2. * (np.asarray(self.sum_diff_f['upsilon_%d_%d' % (p, p)], dtype=np.float64) / T_f)
        res -= 2. * (np.asarray(self.sum_diff_f['phi_%d' % p], dtype=np.float64) / T_f)
        res += 2. * (np.asarray(self.sum_diff_f['psi_%d' % p][1], dtype=np.float64) / T_f)
        res += np.asarray(self.sum_diff_f['upsilonzero_%d' % p], dtype=np.float64) / T_f
        return res

    def recombine_grad_lse_k(self):
        """mhp_stratified_estimator.py:1055-1068."""
        d, k = self.d, self.k
        grad = np.zeros(self.n_param_k)
        grad[0] = self.recombine_diff_lse_mu_k()
        for p in range(d):
            lo, hi = self.interval_map[k][p]
            try:
                block = self._recombine_theta_block(p)
                if block.shape != (hi - lo,):
                    raise ValueError
                grad[lo:hi] = block
            except (ValueError, TypeError, IndexError):
                # (result blocks of another shape, e.g. per-stratum logs: the scalar form)
                for ix in range(self.matrix_n_param[k][p]):
                    ix_param = lo + ix
                    r = self.ix_map[k][ix_param]['par']
                    grad[ix_param] = self.recombine_diff_lse_theta_kpr(p, r)
        return grad

    # ---- the reference's per-term entry points (mhp_exact_estimator.py:543-707,
    # mhp_stratified_estimator.py:1081-1331).  compute_objective batches all terms into one device
    # call; these evaluate one term each for callers that use them directly.  The parameters of the
    # kernels a term does not involve are placeholders (ones): they only enter terms that are not
    # stored.
    def _x_with(self, blocks, mu_k=1.):
        x = np.ones(self.n_param_k)
        x[0] = mu_k
        for i, x_ki in blocks:
            lo, hi = self.interval_map[self.k][i]
            x[lo:hi] = np.asarray(x_ki, dtype=np.float64)
        return x

    def compute_psi_sum(self, i, x_ki, mu_k, **kwargs):
        self._compute_terms(self._x_with([(i, x_ki)], mu_k), [('psi_%d' % i, KIND_PSI, i, i)], **kwargs)

    def compute_upsilonzero_sum(self, i, x_ki, **kwargs):
        self._compute_terms(self._x_with([(i, x_ki)]), [('upsilonzero_%d' % i, KIND_UPSILONZERO, i, i)], **kwargs)

    def compute_phi_sum(self, i, x_ki, **kwargs):
        self._compute_terms(self._x_with([(i, x_ki)]), [('phi_%d' % i, KIND_PHI, i, i)], **kwargs)

    def compute_cross_upsilon_sum(self, i, j, x_ki, x_kj, **kwargs):
        self._compute_terms(self._x_with([(i, x_ki), (j, x_kj)]),
                            [('upsilon_%d_%d' % (i, j), KIND_UPSILON, i, j)], **kwargs)

    def compute_sim_upsilon_sum(self, i, x_ki, **kwargs):
        self._compute_terms(self._x_with([(i, x_ki)]), [('upsilon_%d_%d' % (i, i), KIND_UPSILON, i, i)], **kwargs)

    # ---- stratifications (mhp_stratified_estimator.py:39-117; the exact estimator of the
    # reference builds them too, mhp_exact_estimator.py:39-117, and writes them with save) ----------
    def set_singlesum_estimation(self, n_exact=None, n_samples_adaptive=None,
                                 nonadaptive_sample_size=None, **kwargs):
        d = self.d
        local = {}
        if n_exact is not None:
            local['n_exact'] = n_exact
        if n_samples_adaptive is not None:
            local['n_samples_adaptive'] = n_samples_adaptive
        if nonadaptive_sample_size is not None:
            local['nonadaptive_sample_size'] = nonadaptive_sample_size
        for key, prefix in (('psi_stratification', 'psi_'),
                            ('upsilonzero_stratification', 'upsilonzero_')):
            given = kwargs.get(key, [None] * d)
            for i in range(d):
                strf = given[i]
                if strf is None:
                    strf = GeneralStratification(n=self.n_events[i], **local)
                self.stratif[prefix + str(i)] = strf

    def set_doublesum_estimation(self, n_samples_adaptive=None, nonadaptive_sample_size=None,
                                 **kwargs):
        d, k = self.d, self.k
        local = {}
        if n_samples_adaptive is not None:
            local['n_samples_adaptive'] = n_samples_adaptive
        if nonadaptive_sample_size is not None:
            local['nonadaptive_sample_size'] = nonadaptive_sample_size
        given = kwargs.get('upsilon_stratification', [[None] * d for _ in range(d)])
        for i, j in itertools.product(range(d), range(d)):
            strf = given[i][j]
            if strf is None:
                strf = DoubleIxStratification(h_max=self.kappa[j][i][self.n_events[i] - 1] + 1,
                                              **local)
                strf.get_stratification_size(self.lag_sizes, i, j)
            self.stratif['upsilon_%d_%d' % (i, j)] = strf
        given = kwargs.get('phi_stratification', [None] * d)
        for i in range(d):
            strf = given[i]
            if strf is None:
                strf = DoubleIxStratification(h_max=self.kappa[i][k][self.n_events[k] - 1] + 1,
                                              **local)
                strf.get_stratification_size(self.lag_sizes, k, i)
            self.stratif['phi_%d' % i] = strf

    def set_stratification(self, n_exact_single=None, n_samples_adaptive_single=None,
                           nonadaptive_sample_size_single=None, single_strfs=None,
                           n_samples_adaptive_double=None, nonadaptive_sample_size_double=None,
                           double_strfs=None):
        self.stratif = {}
        self.set_singlesum_estimation(n_exact=n_exact_single,
                                      n_samples_adaptive=n_samples_adaptive_single,
                                      nonadaptive_sample_size=nonadaptive_sample_size_single,
                                      **(single_strfs or {}))
        self.set_doublesum_estimation(n_samples_adaptive=n_samples_adaptive_double,
                                      nonadaptive_sample_size=nonadaptive_sample_size_double,
                                      **(double_strfs or {}))

    def get_lse_k_components_log(self):
        """(n_iter x n_terms table, names): what each term added to LSE_k at every iteration
        (mhp_stratified_estimator.py:476-517; needs is_log_total_estimates)."""
        d, T_f = self.d, self.T_f
        weights = [('psi_%d' % i, 2.) for i in range(d)]
        weights += [('upsilon_%d_%d' % (i, j), 2.) for i, j in itertools.product(range(d), range(d))]
        weights += [('phi_%d' % i, -2.) for i in range(d)]
        weights += [('upsilonzero_%d' % i, 1.) for i in range(d)]
        table = np.zeros((self.n_iter, 1 + len(weights)))
        table[:, 0] = self.mu_term_log
        for col, (name, w) in enumerate(weights, start=1):
            table[:, col] = w * (self.sum_f_log[name] / T_f)
        return table, ['mu_term'] + [name for name, _ in weights]

    def get_log(self):
        logs = {}
        if self.is_log_lse:
            logs['lse'] = self.logged_lse
        if self.is_log_ixs:
            logs['samples'] = copy.deepcopy(self.logged_ixs)
        if self.is_log_allocs:
            logs['allocs'] = copy.deepcopy(self.logged_allocs)
        return logs

    # ---- serialization: the reference's file layout (mhp_stratified_estimator.py:1509-1689,
    # mhp_exact_estimator.py:793-935), so that ``MHP.save`` / ``MHP.load_fit_estimators``
    # (mhp.py:2428-2437, :2489-2506) work with these estimators and files written by either
    # implementation load into the other.  Device handles, the model and the path are not saved;
    # ``load`` takes them as arguments like the reference's.
    _BASIC_ATTRS = ('is_grad_target', 'is_log_ixs', 'is_log_allocs', 'is_log_total_estimates',
                    'is_log_strata_estimates', 'is_log_lse', 'k', 'n_iter')
    _STATE_ATTRS = ('t', 'mu_term', 'diff_mu_term', 'sum_f', 'sum_diff_f')
    _TOTAL_LOG_ATTRS = ('mu_term_log', 'diff_mu_term_log', 'sum_f_log', 'sum_diff_f_log')
    _STRATA_LOG_ATTRS = ()
    _STRATIFICATIONS_OPTIONAL = False   # exact estimators of marked models keep none

    @staticmethod
    def _part(file, suffix):
        return file[:-7] + '_' + suffix + '.pickle' if file.endswith('.pickle') else file + '_' + suffix

    def _saved_stratifications(self):
        return self.stratif

    def save(self, file, **kwargs):
        def dump(suffix, obj):
            with open(self._part(file, suffix), 'wb', **kwargs) as f:
                pickle.dump(obj, f)
        dump('basic_attr', {a: getattr(self, a) for a in self._BASIC_ATTRS})
        stratif = self._saved_stratifications()
        for name in self.list_log_names:
            if name in stratif or not self._STRATIFICATIONS_OPTIONAL:
                stratif[name].save(self._part(file, 'strf_' + name))
        dump('current_state', {a: getattr(self, a) for a in self._STATE_ATTRS})
        general = {}
        if self.is_log_ixs:
            general['logged_ixs'] = self.logged_ixs
        if self.is_log_allocs:
            general['logged_allocs'] = self.logged_allocs
        if self.is_log_lse:
            general['logged_lse'] = self.logged_lse
        if general:
            dump('general_logs', general)
        estimates = {}
        if self.is_log_total_estimates:
            estimates.update({a: getattr(self, a) for a in self._TOTAL_LOG_ATTRS})
        if self.is_log_strata_estimates:
            estimates.update({a: getattr(self, a) for a in self._STRATA_LOG_ATTRS})
        # written even when empty: upstream's load opens this file unconditionally (:1680-1689), so its
        # own files do not load back unless an estimate log was kept
        dump('estimates_logs', estimates)

    def load(self, file, model, process_path, **kwargs):
        def read(suffix, required=True):
            try:
                with open(self._part(file, suffix), 'rb') as f:
                    return pickle.load(f)
            except FileNotFoundError:
                if required:
                    raise
                return {}
        for attr, val in read('basic_attr').items():
            setattr(self, attr, val)
        self.initialize_model_data(model, process_path)
        self.stratif = {}
        for names, cls in ((self.get_singlesum_lognames(), GeneralStratification),
                           (self.get_doublesum_lognames(), DoubleIxStratification)):
            for name in names:
                if self._STRATIFICATIONS_OPTIONAL and not os.path.exists(self._part(file, 'strf_' + name)):
                    continue
                strf = cls()
                strf.load(self._part(file, 'strf_' + name))
                self.stratif[name] = strf
        self.list_log_names = self.get_list_log_names()
        self.make_param_indexing()
        for attr, val in read('current_state').items():
            setattr(self, attr, val)
        general = read('general_logs', required=False)
        names = self.list_log_names
        self.logged_ixs = (general['logged_ixs'] if self.is_log_ixs else
                           {name: {'adaptive': None, 'nonadaptive': None} for name in names})
        self.logged_allocs = (general['logged_allocs'] if self.is_log_allocs else
                              {name: None for name in names})
        if self.is_log_lse:
            self.logged_lse = general['logged_lse']
        # upstream opens the estimates file unconditionally and fails when no estimate log was
        # kept (:1680-1689); here the file is optional like the general logs
        for attr, val in read('estimates_logs', required=False).items():
            setattr(self, attr, val)


class MHPExactEstimB200(_B200Estimator):
    """Exact O(N x window) LSE_k and gradient (MHPExactEstim).

    ``reference_recombine=True`` (default) reproduces the reference's gradient
    assembly verbatim, including the loop of mhp_exact_estimator.py:526-529 that
    overwrites every kernel entry with the p = d-1 combination (SURVEY quirk 1:
    wrong for d >= 2); ``False`` assembles with the stratified estimator's loop,
    which matches finite differences.  The per-term sums are identical either way.
    """

    def __init__(self, *args, reference_recombine=True, **kwargs):
        super().__init__(*args, **kwargs)
        self.reference_recombine = reference_recombine

    def initialize(self, k, n_iter, model, data, n_exact_single=None,
                   n_samples_adaptive_single=None, nonadaptive_sample_size_single=None,
                   single_strfs=None, n_samples_adaptive_double=None,
                   nonadaptive_sample_size_double=None, double_strfs=None):
        # the stratification keywords are accepted for signature compatibility; an
        # exact estimator samples nothing
        self.k = k
        self.n_iter = n_iter
        self.initialize_model_data(model, data)
        self.stratif = {}
        self._strf_kwargs = dict(n_exact_single=n_exact_single,
                                 n_samples_adaptive_single=n_samples_adaptive_single,
                                 nonadaptive_sample_size_single=nonadaptive_sample_size_single,
                                 single_strfs=single_strfs,
                                 n_samples_adaptive_double=n_samples_adaptive_double,
                                 nonadaptive_sample_size_double=nonadaptive_sample_size_double,
                                 double_strfs=double_strfs)
        self.list_log_names = self.get_list_log_names()
        self.make_param_indexing()
        self.initialize_logs()

    def _saved_stratifications(self):
        """The reference's exact estimator carries (unused) stratifications and its ``load`` expects
        their files; they are built here only when the estimator is saved (O(sum of lag ranges))."""
        if not self.stratif:
            self.set_stratification(**getattr(self, '_strf_kwargs', {}))
        return self.stratif

    def initialize_logs(self):
        n_iter = self.n_iter
        self.logged_ixs = {name: {'adaptive': None, 'nonadaptive': None}
                           for name in self.list_log_names}
        self.logged_allocs = {name: None for name in self.list_log_names}
        if self.is_log_lse:
            self.logged_lse = np.zeros(n_iter + 1)
        if self.is_log_total_estimates:
            self.mu_term_log = np.zeros(n_iter)
            self.diff_mu_term_log = np.zeros(n_iter)
            self.sum_f_log = {name: np.zeros(n_iter) for name in self.list_log_names}
            self.sum_diff_f_log = {}
            for name in self.list_log_names:
                n_func, n_param = self.n_func[name], self.n_param[name]
                self.sum_diff_f_log[name] = (
                    np.zeros((n_param, n_iter)) if n_func == 1 else
                    [None, np.zeros((n_param[1], n_iter)), np.zeros((n_param[2], n_iter))])

    def clear_estimates(self):
        self.mu_term = 0.
        self.diff_mu_term = 0.
        for name in self.list_log_names:
            self.sum_f[name] = 0.
            self.sum_diff_f[name] = self.make_null_sumarray(self.n_param[name],
                                                            n_func=self.n_func[name])

    def log_estimates(self):
        if not self.is_log_total_estimates:
            return
        t = self.t
        self.mu_term_log[t] = self.mu_term
        self.diff_mu_term_log[t] = self.diff_mu_term
        for name in self.list_log_names:
            self.sum_f_log[name][t] = self.sum_f[name]
            if self.n_func[name] == 1:
                self.sum_diff_f_log[name][:, t] = self.sum_diff_f[name] + 0.
            else:
                for ix_func in (1, 2):
                    self.sum_diff_f_log[name][ix_func][:, t] = self.sum_diff_f[name][ix_func] + 0.

    def compute_objective(self, x_k, compute_f_sum=False, compute_diff_f_sum=True,
                          count_iter=True, verbose=False):
        """One device call for all terms (mhp_exact_estimator.py:709-761)."""
        self.clear_estimates()
        x = np.asarray(x_k, dtype=np.float64)   # the fit loop passes dtype=object arrays
        self.compute_mu_term(x[0], compute_f_sum=compute_f_sum,
                             compute_diff_f_sum=compute_diff_f_sum)
        if compute_f_sum or compute_diff_f_sum:
            engine = self._bound_engine()
            packed = engine.exact_sums_packed(self.k, x, compute_f_sum, compute_diff_f_sum)
            sum_f, sum_diff_f = unpack_terms(engine.desc, self.k, packed)
            for name in self.list_log_names:
                if compute_f_sum:
                    self.sum_f[name] = sum_f[name]
                if compute_diff_f_sum:
                    self.sum_diff_f[name] = sum_diff_f[name]
        self.log_estimates()
        if count_iter:
            self.t += 1

    def _compute_terms(self, x, terms, compute_f_sum=False, compute_diff_f_sum=True, verbose=False,
                       **kwargs):
        """The reference's per-term methods set the value whenever the derivatives are asked for
        (:557-558).  The device evaluates the whole sub-problem; only ``terms`` are stored."""
        if not (compute_f_sum or compute_diff_f_sum):
            return
        engine = self._bound_engine()
        packed = engine.exact_sums_packed(self.k, x, True, compute_diff_f_sum)
        sum_f, sum_diff_f = unpack_terms(engine.desc, self.k, packed)
        for name, _, _, _ in terms:
            self.sum_f[name] = sum_f[name]
            if compute_diff_f_sum:
                self.sum_diff_f[name] = sum_diff_f[name]

    def recombine_grad_lse_k(self):
        if not self.reference_recombine:
            return super().recombine_grad_lse_k()
        # loop structure of mhp_exact_estimator.py:518-530: for every p ALL kernel entries are
        # written, so what stays is the p = d - 1 combination at every parameter's own index r
        d, k = self.d, self.k
        grad = np.zeros(self.n_param_k)
        grad[0] = self.recombine_diff_lse_mu_k()
        pars = [self.ix_map[k][ix_param]['par'] for ix_param in range(1, self.n_param_k)]
        r_max = max(pars, default=-1)
        if pars and all(r_max < self.matrix_n_param[k][p] for p in range(d)):
            # (every p of the upstream loop evaluates without an IndexError)
            grad[1:] = self._recombine_theta_block(d - 1)[pars]
            return grad
        for p in range(d):
            for ix_param in range(1, self.n_param_k):
                r = self.ix_map[k][ix_param]['par']
                grad[ix_param] = self.recombine_diff_lse_theta_kpr(p, r)
        return grad

    def lse_k_estimate(self, x_k, compute_diff_f_sum=False, count_iter=True, verbose=False,
                       **kwargs):
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


class MHPStratEstimB200(_B200Estimator):
    """Adaptive stratified sampling estimator (ASLSD; MHPStratEstim).

    Per call of ``compute_objective`` the host draws every sample of every term
    with the reference's RNG call sequence (``aslsd_b200.sampling``), one device
    launch evaluates all of them and returns per-stratum sums, the allocation
    statistic and the updated relative allocations; the host then fills the
    reference's result attributes.
    """

    def __init__(self, *args, reference_adaptive_grad=True, saturated_values=True, **kwargs):
        super().__init__(*args, **kwargs)
        # True: samples of psi_i / upsilonzero_i far from T_f carry the value of the model's own
        # closures (bit for bit the reference's), so that the rounding-residue allocations of such
        # strata - hence every later draw - are the reference's (aslsd_b200.saturation)
        self.saturated_values = saturated_values
        # True: verbatim reference behaviour - without grad_alloc the derivative
        # means of the ADAPTIVE strata are never evaluated (the ``if n_func == 1``
        # block of mhp_stratified_estimator.py:607-627 is nested under
        # ``if grad_alloc``), so those strata add zero to the gradient estimate.
        # False: the adaptive strata contribute size * mean like the others.
        self.reference_adaptive_grad = reference_adaptive_grad
        self.vecsum_f_adaptive, self.sum_f_adaptive = {}, {}
        self.vecsum_f_nonadaptive, self.sum_f_nonadaptive = {}, {}
        self.sum_f_exact = {}
        self.vecsum_diff_f_adaptive, self.sum_diff_f_adaptive = {}, {}
        self.vecsum_diff_f_nonadaptive, self.sum_diff_f_nonadaptive = {}, {}
        self.sum_diff_f_exact = {}

    # ---- stratification (mhp_stratified_estimator.py:55-204) ------------------------------
    _STATE_ATTRS = ('t', 'mu_term', 'diff_mu_term', 'sum_f', 'vecsum_f_adaptive', 'sum_f_adaptive',
                    'vecsum_f_nonadaptive', 'sum_f_nonadaptive', 'sum_f_exact', 'sum_diff_f',
                    'vecsum_diff_f_adaptive', 'sum_diff_f_adaptive', 'vecsum_diff_f_nonadaptive',
                    'sum_diff_f_nonadaptive', 'sum_diff_f_exact')
    _STRATA_LOG_ATTRS = ('vecsum_f_adaptive_log', 'sum_f_adaptive_log', 'vecsum_f_nonadaptive_log',
                         'sum_f_nonadaptive_log', 'sum_f_exact_log', 'vecsum_diff_f_adaptive_log',
                         'sum_diff_f_adaptive_log', 'vecsum_diff_f_nonadaptive_log',
                         'sum_diff_f_nonadaptive_log', 'sum_diff_f_exact_log')

    def initialize(self, k, n_iter, model, data, n_exact_single=None,
                   n_samples_adaptive_single=None, nonadaptive_sample_size_single=None,
                   single_strfs=None, n_samples_adaptive_double=None,
                   nonadaptive_sample_size_double=None, double_strfs=None):
        self.k = k
        self.n_iter = n_iter
        self.initialize_model_data(model, data)
        self.set_stratification(n_exact_single=n_exact_single,
                                n_samples_adaptive_single=n_samples_adaptive_single,
                                nonadaptive_sample_size_single=nonadaptive_sample_size_single,
                                single_strfs=single_strfs,
                                n_samples_adaptive_double=n_samples_adaptive_double,
                                nonadaptive_sample_size_double=nonadaptive_sample_size_double,
                                double_strfs=double_strfs)
        self.list_log_names = self.get_list_log_names()
        self.make_param_indexing()
        self.initialize_logs()

    # ---- result containers and logs (:272-474) -------------------------------------------------
    def clear_estimates(self):
        self.mu_term = 0.
        self.diff_mu_term = 0.
        for name in self.list_log_names:
            strf = self.stratif[name]
            n_func, n_param = self.n_func[name], self.n_param[name]
            self.sum_f[name] = 0.
            self.vecsum_f_adaptive[name] = np.zeros(strf.n_adaptive_strata)
            self.sum_f_adaptive[name] = 0.
            self.vecsum_f_nonadaptive[name] = np.zeros(strf.n_nonadaptive_strata)
            self.sum_f_nonadaptive[name] = 0.
            self.sum_f_exact[name] = 0.
            null = self.make_null_sumarray
            self.sum_diff_f[name] = null(n_param, n_func=n_func)
            self.vecsum_diff_f_adaptive[name] = null(n_param, n_func=n_func,
                                                     n_strata=strf.n_adaptive_strata)
            self.sum_diff_f_adaptive[name] = null(n_param, n_func=n_func)
            self.vecsum_diff_f_nonadaptive[name] = null(n_param, n_func=n_func,
                                                        n_strata=strf.n_nonadaptive_strata)
            self.sum_diff_f_nonadaptive[name] = null(n_param, n_func=n_func)
            self.sum_diff_f_exact[name] = null(n_param, n_func=n_func)

    _STRATA_LOGS = ('vecsum_f_adaptive', 'sum_f_adaptive', 'vecsum_f_nonadaptive',
                    'sum_f_nonadaptive', 'sum_f_exact')
    _STRATA_DIFF_LOGS = ('vecsum_diff_f_adaptive', 'sum_diff_f_adaptive',
                         'vecsum_diff_f_nonadaptive', 'sum_diff_f_nonadaptive',
                         'sum_diff_f_exact')

    def _null_log(self, name, n_strata, n_iter):
        n_func, n_param = self.n_func[name], self.n_param[name]

        def zeros(n):
            return np.zeros((n, n_iter)) if n_strata is None else np.zeros((n, n_strata, n_iter))
        return zeros(n_param) if n_func == 1 else [None, zeros(n_param[1]), zeros(n_param[2])]

    def initialize_logs(self):
        n_iter = self.n_iter
        names = self.list_log_names
        self.logged_ixs = {name: {'adaptive': [None] * n_iter if self.is_log_ixs else None,
                                  'nonadaptive': [None] * n_iter if self.is_log_ixs else None}
                           for name in names}
        self.logged_allocs = {name: [None] * n_iter if self.is_log_allocs else None
                              for name in names}
        if self.is_log_lse:
            self.logged_lse = np.zeros(n_iter + 1)
        if self.is_log_total_estimates:
            self.mu_term_log = np.zeros(n_iter)
            self.diff_mu_term_log = np.zeros(n_iter)
            self.sum_f_log = {name: np.zeros(n_iter) for name in names}
            self.sum_diff_f_log = {name: self._null_log(name, None, n_iter) for name in names}
        if self.is_log_strata_estimates:
            for attr in self._STRATA_LOGS + self._STRATA_DIFF_LOGS:
                setattr(self, attr + '_log', {})
            for name in names:
                strf = self.stratif[name]
                n_a, n_n = strf.n_adaptive_strata, strf.n_nonadaptive_strata
                self.vecsum_f_adaptive_log[name] = np.zeros((n_a, n_iter))
                self.sum_f_adaptive_log[name] = np.zeros(n_iter)
                self.vecsum_f_nonadaptive_log[name] = np.zeros((n_n, n_iter))
                self.sum_f_nonadaptive_log[name] = np.zeros(n_iter)
                self.sum_f_exact_log[name] = np.zeros(n_iter)
                self.vecsum_diff_f_adaptive_log[name] = self._null_log(name, n_a, n_iter)
                self.sum_diff_f_adaptive_log[name] = self._null_log(name, None, n_iter)
                self.vecsum_diff_f_nonadaptive_log[name] = self._null_log(name, n_n, n_iter)
                self.sum_diff_f_nonadaptive_log[name] = self._null_log(name, None, n_iter)
                self.sum_diff_f_exact_log[name] = self._null_log(name, None, n_iter)

    def _log_diff(self, attr, name, t):
        logs, val = getattr(self, attr + '_log')[name], getattr(self, attr)[name]
        if self.n_func[name] == 1:
            logs[..., t] = val + 0.
        else:
            for ix_func in (1, 2):
                logs[ix_func][..., t] = val[ix_func] + 0.

    def log_estimates(self):
        t = self.t
        if self.is_log_total_estimates:
            self.mu_term_log[t] = self.mu_term
            self.diff_mu_term_log[t] = self.diff_mu_term
            for name in self.list_log_names:
                self.sum_f_log[name][t] = self.sum_f[name]
                self._log_diff('sum_diff_f', name, t)
        if self.is_log_strata_estimates:
            for name in self.list_log_names:
                for attr in self._STRATA_LOGS:
                    getattr(self, attr + '_log')[name][..., t] = getattr(self, attr)[name] + 0.
                for attr in self._STRATA_DIFF_LOGS:
                    self._log_diff(attr, name, t)

    # ---- sampling plan of one term -------------------------------------------------------------
    def _draw(self, kind, i, j, strata, n_samples, rng):
        """Samples of the given strata: (per-stratum list for the index log or None, all m in one
        array, all n in one array - None for single-index terms -, samples per stratum); RNG
        consumption as in the compute_*_sum closures."""
        if kind in (KIND_PSI, KIND_UPSILONZERO):
            m, counts = sampling.sample_single_ix_stratified(strata, n_samples, rng, flat=True)
            n = None
        else:
            tgt, src = (self.k, i) if kind == KIND_PHI else (i, j)
            m, n, counts = sampling.sample_double_ix_stratified(tgt, src, strata, n_samples, self.kappa,
                                                                self.varpi, self.n_events, rng, flat=True)
        raw = None
        if self.is_log_ixs:     # the reference logs one array (single) / one {'m', 'n'} dict (double) per stratum
            offs = np.concatenate([[0], np.cumsum(counts)])
            if n is None:
                raw = [m[offs[q]:offs[q + 1]] for q in range(len(counts))]
            else:
                raw = [{'m': m[offs[q]:offs[q + 1]], 'n': n[offs[q]:offs[q + 1]]} for q in range(len(counts))]
        return raw, m, n, counts

    def compute_objective(self, x_k, compute_f_sum=False, compute_diff_f_sum=True,
                          grad_alloc=True, rng=None, seed=None, count_iter=True):
        """mhp_stratified_estimator.py:1333-1386 with estimate_sum (:902-971) of
        every term batched into one device launch."""
        self.clear_estimates()
        if rng is None:
            rng = np.random.default_rng(seed)
        x = np.asarray(x_k, dtype=np.float64)
        self.compute_mu_term(x[0], compute_f_sum=compute_f_sum,
                             compute_diff_f_sum=compute_diff_f_sum)
        self.last_alloc_stat = {}   # name -> (mean, std, n_samples) of the allocation statistic per adaptive stratum
        self._estimate_terms(x, self.term_list(), compute_f_sum, compute_diff_f_sum, grad_alloc, rng)
        self.log_estimates()
        if count_iter:
            self.t += 1

    def _estimate_terms(self, x, terms, compute_f_sum, compute_diff_f_sum, grad_alloc, rng):
        """estimate_sum (:902-971) of the given terms: samples drawn term after term, one device
        launch for all of them, results into the reference's attributes."""
        t_iter = self.t
        # ---- 1. draw every sample, term after term (host RNG, reference order) ----------
        plans = []
        dev_terms, all_m, all_n, all_counts, sizes, defaults = [], [], [], [], [], []

        def add(m, n, counts, size_vec, default_vec):
            all_m.append(m)
            all_n.append(n if n is not None else np.zeros(len(m), dtype=np.int64))
            all_counts.append(np.asarray(counts, dtype=np.int64))
            sizes.append(size_vec)
            defaults.append(default_vec)
        for name, kind, i, j in terms:
            strf = self.stratif[name]
            plan = {'name': name, 'blocks': [], 'counts': []}   # blocks: (domain, n_strata)
            if strf.n_adaptive_strata > 0:
                n_a = strf.n_adaptive_strata
                rel_alloc = (strf.ema_weight * strf.adaptive_rel_alloc
                             + (1. - strf.ema_weight) * strf.guess_rel_alloc)      # :566-567
                strf.guess_rel_alloc = copy.deepcopy(rel_alloc)                    # :570
                n_samples = [min(strf.adaptive_strata_sizes[s],
                                 max(1, int(rel_alloc[s] * (n_a + strf.n_samples_adaptive[0]))))
                             for s in range(n_a)]                                  # :587
                raw, m, n, counts = self._draw(kind, i, j, strf.adaptive_strata, n_samples, rng)
                if self.is_log_ixs:
                    self.logged_ixs[name]['adaptive'][t_iter] = [[s] for s in raw]
                add(m, n, counts, np.asarray(strf.adaptive_strata_sizes, dtype=np.float64),
                    np.asarray(rel_alloc, dtype=np.float64))
                plan['blocks'].append(('adaptive', n_a))
                plan['counts'].append(counts)
            if strf.n_nonadaptive_strata > 0:
                raw, m, n, counts = self._draw(kind, i, j, strf.nonadaptive_strata,
                                               strf.nonadaptive_abs_alloc, rng)
                if self.is_log_ixs:
                    self.logged_ixs[name]['nonadaptive'][t_iter] = raw
                add(m, n, counts, np.zeros(strf.n_nonadaptive_strata), np.zeros(strf.n_nonadaptive_strata))
                plan['blocks'].append(('nonadaptive', strf.n_nonadaptive_strata))
                plan['counts'].append(counts)
            if strf.n_exact > 0:
                m = np.asarray(strf.exact_domain)
                add(m, None, [len(m)], np.zeros(1), np.zeros(1))
                plan['blocks'].append(('exact', 1))
                plan['counts'].append(np.array([len(m)], dtype=np.int64))
            n_chunks = sum(n_strata for _, n_strata in plan['blocks'])
            if n_chunks:
                sat = None
                if kind in (KIND_PSI, KIND_UPSILONZERO) and self.saturated_values:
                    # the reference's own value of the samples far from T_f (aslsd_b200.saturation)
                    sat = saturation.term_saturation(self._model, self.desc, self.k, i, kind, x)
                dev_terms.append((kind, i, j, n_chunks) + (sat if sat else ()))
            plans.append(plan)
        # ---- 2. one launch ----------------------------------------------------------------------
        if dev_terms:
            engine = self._bound_engine()
            ix_m = np.concatenate(all_m)
            ix_n = np.concatenate(all_n)
            offsets = np.concatenate([[0], np.cumsum(np.concatenate(all_counts))])
            sums, out_off, stat_mean, stat_std, rel = engine.sampled_sums(
                self.k, x, dev_terms, offsets, ix_m, ix_n, strata_sizes=np.concatenate(sizes),
                default_alloc=np.concatenate(defaults),
                compute_diff_f_sum=bool(compute_diff_f_sum or grad_alloc), grad_alloc=bool(grad_alloc))
        # ---- 3. fill the reference's result attributes -----------------------------------------------
        s_ix = 0
        if not hasattr(self, 'last_alloc_stat'):
            self.last_alloc_stat = {}
        for plan in plans:
            name = plan['name']
            strf = self.stratif[name]
            n_func, n_param = self.n_func[name], self.n_param[name]
            for (domain, n_strata), counts in zip(plan['blocks'], plan['counts']):
                # which sums each domain produces (:549-554, :795-796, :872-873)
                do_f, do_diff = compute_f_sum, compute_diff_f_sum
                if domain == 'adaptive':
                    if grad_alloc:
                        do_diff = True
                    else:
                        do_f = True
                if do_diff and n_func > 1:
                    do_f = True
                if domain == 'adaptive' and not grad_alloc and self.reference_adaptive_grad:
                    do_diff = False   # reference quirk: see __init__
                # the strata of a term are contiguous and have the same number of outputs
                block = sums[out_off[s_ix]:out_off[s_ix + n_strata]].reshape(n_strata, -1)   # (n_strata, n_out)
                counts = np.asarray(counts, dtype=np.float64)
                if domain == 'exact':
                    est = block                                          # plain sums (:883)
                else:
                    strata_sizes = np.asarray(strf.adaptive_strata_sizes if domain == 'adaptive'
                                              else strf.nonadaptive_strata_sizes, dtype=np.float64)
                    est = (block / counts[:, None]) * strata_sizes[:, None]   # size * mean
                self._store(name, domain, est, do_f, do_diff)
                if domain == 'adaptive':
                    self.last_alloc_stat[name] = (stat_mean[s_ix:s_ix + n_strata].copy(),
                                                  stat_std[s_ix:s_ix + n_strata].copy(), counts)
                    strf.adaptive_rel_alloc = rel[s_ix:s_ix + n_strata].copy()     # :740
                    if self.is_log_allocs:
                        self.logged_allocs[name][t_iter] = strf.adaptive_rel_alloc.copy()
                s_ix += n_strata
            # totals (:947-967)
            if compute_f_sum:
                self.sum_f[name] = (self.sum_f_adaptive[name] + self.sum_f_nonadaptive[name]
                                    + self.sum_f_exact[name])
            if compute_diff_f_sum:
                parts = (self.sum_diff_f_adaptive[name], self.sum_diff_f_nonadaptive[name],
                         self.sum_diff_f_exact[name])
                if n_func == 1:
                    self.sum_diff_f[name] = parts[0] + parts[1] + parts[2]
                else:
                    for ix_func in (1, 2):
                        self.sum_diff_f[name][ix_func] = (parts[0][ix_func] + parts[1][ix_func]
                                                          + parts[2][ix_func])

    def _compute_terms(self, x, terms, compute_f_sum=False, compute_diff_f_sum=True, grad_alloc=True,
                       rng=None, seed=None, **kwargs):
        if rng is None:
            rng = np.random.default_rng(seed)
        self._estimate_terms(x, terms, compute_f_sum, compute_diff_f_sum, grad_alloc, rng)

    def _store(self, name, domain, est, do_f, do_diff):
        """est[s, c]: estimate of the sum over stratum s of output c (0 = f, then
        the derivatives of ix_func 1, then those of ix_func 2)."""
        n_func, n_param = self.n_func[name], self.n_param[name]
        if do_f:
            if domain == 'exact':
                self.sum_f_exact[name] = est[0, 0]
            else:
                vec = est[:, 0].copy()
                getattr(self, 'vecsum_f_' + domain)[name] = vec
                getattr(self, 'sum_f_' + domain)[name] = vec.sum()
        if not do_diff:
            return
        if n_func == 1:
            splits = [(None, 1, 1 + n_param)]
        else:
            splits = [(1, 1, 1 + n_param[1]), (2, 1 + n_param[1], 1 + n_param[1] + n_param[2])]
        totals = getattr(self, 'sum_diff_f_' + domain)
        for ix_func, a, b in splits:
            per_param = est[:, a:b].T.copy()             # (n_param, n_strata)
            if domain != 'exact':
                vecs = getattr(self, 'vecsum_diff_f_' + domain)
                if ix_func is None:
                    vecs[name] = per_param
                else:
                    vecs[name][ix_func] = per_param
            total = per_param.sum(axis=1)
            if ix_func is None:
                totals[name] = total
            else:
                totals[name][ix_func] = total

    # ---- LSE and gradient (:1391-1427) --------------------------------------------------------------
    def lse_k_estimate(self, x_k, compute_diff_f_sum=False, grad_alloc=True, rng=None,
                       seed=None, count_iter=True, **kwargs):
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
