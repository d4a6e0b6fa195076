"""Oracle for the exact LSE sums of a marked model (MTLH) sub-problem k (test infrastructure).

Restates ``MTLHExactEstim.compute_objective`` and its ``compute_*_sum`` methods
(optimize/estimators/mtlh_exact_estimator.py:674-1030) with the reference's own loop
structure, and the LSE / gradient assembly (``recombine_*`` :509-675), for the model
families the B200 engine serves: Exponential / Gaussian basis kernels, constant
baselines (functionals/baselines/basis_baselines/basis_baseline_constant.py:30-65) and baselines
constant on time intervals (basis_baseline_piececonst.py:108-300, ``model['baselines'][k] =
{'type': 'piececonst', 'beta': breakpoints}``; its K is restated literally, with the reference's
last-interval term psi(t) - psi(beta_G) as written there), Linear and
Constant impact functions (functionals/impact_functions/basis_impacts/basis_impact_linear.py:
36-52, basis_impact_constant.py:33-40).

Model description (plain data): ``{'d', 'kernels', 'impacts'}`` with ``kernels[k][i]`` as in
oracle/exact.py and ``impacts[k][i]`` = ``{'type': 'linear', 'mark_dim': q}`` or
``{'type': 'constant'}``.  ``x_k = [b_k, params(phi_k0), ..., params(I_k0), ...]``
(models/hawkes/linear/mtlh.py:595-635: baseline, then kernels, then impacts).

Late binding (models/hawkes/linear/mtlh.py:1028-1077): as in MHP, every ``upsilon`` /
``diff_cross_upsilon`` / ``K`` / ``diff_K`` closure ends up bound to the objects of the last loop
iteration - K[d-1][d-1] and baseline d-1; with one kernel structure and constant baselines
everywhere this only matters for which structure is read (``ups_kernel``).
"""
import numpy as np

from oracle import functionals as F
from oracle.exact import ups_kernel


def impact_n_param(spec):
    return spec.get('mark_dim', 1) if spec['type'] == 'linear' else 0


def impact(spec, xi, params):
    """basis_impact_linear.py:36-41 / basis_impact_constant.py:33-37"""
    if spec['type'] == 'linear':
        return 1. + np.sum(xi * np.asarray(params, dtype=float), axis=xi.ndim - 1)
    return np.ones(len(xi))


def diff_impact(spec, xi, ix_diff, params):
    """basis_impact_linear.py:43-52"""
    return xi[:, ix_diff]


def baseline_beta(model, k):
    """Breakpoints of the baseline of sub-problem k as PieceConstBaseline.process_beta keeps them
    (basis_baseline_piececonst.py:59-67), or None for a constant baseline."""
    spec = model.get('baselines', [{'type': 'constant'}] * model['d'])[k]
    if spec['type'] == 'constant':
        return None
    q = np.array(sorted(set(float(v) for v in spec['beta'])))
    q = q[q >= 0.]
    if q[0] != 0.:
        q = np.insert(q, 0, 0.)
    if q[-1] != np.inf:
        q = np.append(q, np.inf)
    return q


def piececonst_K(kernel, beta, t, s, x_ker, x_mu, diff=None):
    """PieceConstBaseline.make_K (basis_baseline_piececonst.py:172-203) summed over the bases of the
    kernel (basis_baseline.py:438-449), array branch; with ``diff = (ix_func, ix_diff)`` make_diff_K
    (:211-296): ix_func 1 the kernel parameter ix_diff, ix_func 2 the baseline value ix_diff."""
    def g(v):
        return np.digitize(v, beta) - 1
    if diff is not None and diff[0] == 1:
        def psi(v):
            return F.diff_psi(kernel, v, diff[1], x_ker)
    else:
        def psi(v):
            return F.psi(kernel, v, x_ker)
    g_s, g_ts = g(s), g(t + s)
    res = np.zeros(len(t))
    if diff is None or diff[0] == 1:
        same = g_s == g_ts
        res[same] = x_mu[g_s[same]] * psi(t[same])
        other = g_s < g_ts
        init = x_mu[g_s[other]] * psi(beta[g_s[other] + 1] - s[other])
        fin = x_mu[g_ts[other]] * (psi(t[other]) - psi(beta[g_ts[other]]))
        res[other] = init + fin
        for ix in np.where(g_s <= g_ts - 2)[0]:
            for j in range(g_s[ix] + 1, g_ts[ix]):
                res[ix] += x_mu[j] * (psi(beta[j + 1] - s[ix:ix + 1]) - psi(beta[j] - s[ix:ix + 1]))[0]
        return res
    ix = diff[1]
    diff_g = g_ts - g_s
    sel = (diff_g == 0) & (g_s == ix)
    res[sel] = psi(t[sel])
    sel = (diff_g > 0) & (g_s == ix)
    res[sel] = psi(beta[ix + 1] - s[sel])
    sel = (diff_g > 0) & (g_ts == ix)
    res[sel] = psi(t[sel]) - psi(beta[ix])
    sel = (diff_g > 1) & (g_s < ix) & (g_ts > ix)
    res[sel] = psi(beta[ix + 1] - s[sel]) - psi(beta[ix] - s[sel])
    return res


def layout(model, k):
    """(n_mu, ker intervals, imp intervals) of x_k (mtlh.py:405-575)."""
    d = model['d']
    beta = baseline_beta(model, k)
    n_mu = 1 if beta is None else len(beta) - 1
    pos = n_mu
    ker, imp = [], []
    for i in range(d):
        n = F.kernel_n_param(model['kernels'][k][i])
        ker.append((pos, pos + n))
        pos += n
    for i in range(d):
        n = impact_n_param(model['impacts'][k][i])
        imp.append((pos, pos + n))
        pos += n
    return n_mu, ker, imp


def n_param_k(model, k):
    return layout(model, k)[2][-1][1]


def _null_sums(model, k):
    d = model['d']
    n_mu, ker, imp = layout(model, k)
    nk = [b - a for a, b in ker]
    ni = [b - a for a, b in imp]
    sum_f = {'mu': 0.}
    sum_diff_f = {'mu': [np.zeros(n_mu)]}
    for i in range(d):
        for name in ('phi_%d' % i, 'upsilonzero_%d' % i):
            sum_f[name] = 0.
            sum_diff_f[name] = [np.zeros(nk[i]), np.zeros(ni[i])]
        name = 'K_%d' % i
        sum_f[name] = 0.
        sum_diff_f[name] = [[None, np.zeros(nk[i]), np.zeros(n_mu)], np.zeros(ni[i])]
    for i in range(d):
        for j in range(d):
            name = 'upsilon_%d_%d' % (i, j)
            sum_f[name] = 0.
            if i == j:
                sum_diff_f[name] = [np.zeros(nk[i]), np.zeros(ni[i])]
            else:
                sum_diff_f[name] = [[None, np.zeros(nk[i]), np.zeros(nk[j])],
                                    [None, np.zeros(ni[i]), np.zeros(ni[j])]]
    return sum_f, sum_diff_f


def exact_sums_literal(model, k, x_k, times, marks, T_f):
    """Returns (M_term, diff_M_term, sum_f, sum_diff_f) like the attributes the reference's
    estimator leaves behind (compute_f_sum and compute_diff_f_sum both on)."""
    x_k = np.asarray(x_k, dtype=float)
    d = model['d']
    K = model['kernels']
    I = model['impacts']
    KU = ups_kernel(model)
    n_mu, ker, imp = layout(model, k)
    beta = baseline_beta(model, k)
    b = x_k[0]
    x_mu = x_k[:n_mu]
    x_ker = [x_k[a:c] for a, c in ker]
    x_imp = [x_k[a:c] for a, c in imp]
    sum_f, sum_diff_f = _null_sums(model, k)
    if beta is None:
        # M term, mu sum (:674-705; constant baseline: M = b^2, mu = b)
        M_term = b ** 2
        diff_M_term = np.array([2. * b])
        sum_f['mu'] = b * len(times[k])
        sum_diff_f['mu'][0][0] = float(len(times[k]))
    else:
        # PieceConstBaseline.make_M / make_diff_M (scalar branch, :113-170), make_mu / make_diff_mu (:108-111)
        sizes = beta[1:] - beta[:-1]
        g_T = int(np.digitize(T_f, beta) - 1)
        M_term = x_mu[g_T] ** 2 * (T_f - beta[g_T])
        if g_T > 0:
            M_term += np.cumsum(x_mu ** 2 * sizes)[g_T - 1]
        M_term /= T_f
        diff_M_term = np.zeros(n_mu)
        for r in range(n_mu):
            if g_T > r:
                diff_M_term[r] = 2. * x_mu[r] * (beta[r + 1] - beta[r]) / T_f
            elif g_T == r:
                diff_M_term[r] = 2. * x_mu[r] * (T_f - beta[r]) / T_f
        tk = times[k]
        sum_f['mu'] = np.sum(x_mu[np.digitize(tk, beta) - 1])
        for r in range(n_mu):
            sum_diff_f['mu'][0][r] = float(np.sum((tk >= beta[r]) & (tk <= beta[r + 1])))
    w = [impact(I[k][i], marks[i], x_imp[i]) for i in range(d)]
    dw = [[diff_impact(I[k][i], marks[i], r, x_imp[i]) for r in range(len(x_imp[i]))]
          for i in range(d)]
    for i in range(d):
        n_i = len(x_ker[i])
        t = T_f - times[i]
        name = 'K_%d' % i
        if beta is None:
            # K term (:707-753): K = b psi for a constant baseline (basis_baseline_constant.py:52-62)
            psi_vals = F.psi(KU, t, x_ker[i])
            sum_f[name] = np.sum(b * psi_vals * w[i])
            for p in range(n_i):
                sum_diff_f[name][0][1][p] = np.sum(b * F.diff_psi(KU, t, p, x_ker[i]) * w[i])
            sum_diff_f[name][0][2][0] = np.sum(psi_vals * w[i])
            for r in range(len(x_imp[i])):
                sum_diff_f[name][1][r] = np.sum(dw[i][r] * (b * psi_vals))
        else:
            k_vals = piececonst_K(KU, beta, t, times[i], x_ker[i], x_mu)
            sum_f[name] = np.sum(k_vals * w[i])
            for p in range(n_i):
                sum_diff_f[name][0][1][p] = np.sum(piececonst_K(KU, beta, t, times[i], x_ker[i], x_mu, (1, p)) * w[i])
            for r in range(n_mu):
                sum_diff_f[name][0][2][r] = np.sum(piececonst_K(KU, beta, t, times[i], x_ker[i], x_mu, (2, r)) * w[i])
            for r in range(len(x_imp[i])):
                sum_diff_f[name][1][r] = np.sum(dw[i][r] * k_vals)
        # upsilon at lag zero (:755-805)
        name = 'upsilonzero_%d' % i
        ups0 = F.upsilon(KU, KU, t, 0., x_ker[i], x_ker[i])
        sum_f[name] = np.sum(ups0 * w[i] ** 2)
        for p in range(n_i):
            sum_diff_f[name][0][p] = np.sum(F.diff_sim_upsilon(K[k][i], t, 0., p, x_ker[i]) * w[i] ** 2)
        for r in range(len(x_imp[i])):
            sum_diff_f[name][1][r] = np.sum(ups0 * 2. * w[i] * dw[i][r])
        # phi term (:807-853): targets of type k, sources of type i weighted by their impact
        name = 'phi_%d' % i
        tk, ti = times[k], times[i]
        for m in range(len(tk)):
            hi = m if i == k else np.searchsorted(ti, tk[m], side='left')
            if hi == 0:
                continue
            s = tk[m] - ti[:hi]
            phi_vals = F.phi(K[k][i], s, x_ker[i])
            sum_f[name] += np.sum(phi_vals * w[i][:hi])
            for p in range(n_i):
                sum_diff_f[name][0][p] += np.sum(F.diff_phi(K[k][i], s, p, x_ker[i]) * w[i][:hi])
            for r in range(len(x_imp[i])):
                sum_diff_f[name][1][r] += np.sum(phi_vals * dw[i][r][:hi])
        # same-type upsilon (:922-971)
        name = 'upsilon_%d_%d' % (i, i)
        for m in range(1, len(ti)):
            s = ti[m] - ti[:m]
            tt = (T_f - ti[m]) * np.ones(m)
            ups = F.upsilon(KU, KU, tt, s, x_ker[i], x_ker[i])
            loc = w[i][m] * w[i][:m]
            sum_f[name] += np.sum(ups * loc)
            for p in range(n_i):
                sum_diff_f[name][0][p] += np.sum(F.diff_sim_upsilon(K[k][i], tt, s, p, x_ker[i]) * loc)
            for r in range(len(x_imp[i])):
                term = dw[i][r][m] * w[i][:m] + w[i][m] * dw[i][r][:m]
                sum_diff_f[name][1][r] += np.sum(ups * term)
    # cross upsilon (:855-920)
    for i in range(d):
        for j in range(d):
            if i == j:
                continue
            name = 'upsilon_%d_%d' % (i, j)
            ti, tj = times[i], times[j]
            for m in range(len(ti)):
                hi = np.searchsorted(tj, ti[m], side='left')
                if hi == 0:
                    continue
                s = ti[m] - tj[:hi]
                tt = (T_f - ti[m]) * np.ones(hi)
                ups = F.upsilon(KU, KU, tt, s, x_ker[i], x_ker[j])
                loc = w[i][m] * w[j][:hi]
                sum_f[name] += np.sum(ups * loc)
                for ix_func, q in ((1, i), (2, j)):
                    for p in range(len(x_ker[q])):
                        sum_diff_f[name][0][ix_func][p] += np.sum(
                            F.diff_cross_upsilon(KU, KU, tt, s, ix_func, p, x_ker[i], x_ker[j]) * loc)
                for r in range(len(x_imp[i])):
                    sum_diff_f[name][1][1][r] += np.sum(ups * (dw[i][r][m] * w[j][:hi]))
                for r in range(len(x_imp[j])):
                    sum_diff_f[name][1][2][r] += np.sum(ups * (w[i][m] * dw[j][r][:hi]))
    return M_term, diff_M_term, sum_f, sum_diff_f


def recombine_lse_k(model, k, M_term, sum_f, T_f):
    """mtlh_exact_estimator.py:509-534"""
    d = model['d']
    res = M_term - 2. * (sum_f['mu'] / T_f)
    for i in range(d):
        res += 2. * (sum_f['K_%d' % i] / T_f)
    for i in range(d):
        for j in range(d):
            res += 2. * (sum_f['upsilon_%d_%d' % (i, j)] / T_f)
    for j in range(d):
        res -= 2. * (sum_f['phi_%d' % j] / T_f)
    for i in range(d):
        res += sum_f['upsilonzero_%d' % i] / T_f
    return res


def recombine_grad_lse_k(model, k, diff_M_term, sum_diff_f, T_f):
    """mtlh_exact_estimator.py:536-670"""
    d = model['d']
    n_mu, ker, imp = layout(model, k)
    grad = np.zeros(imp[-1][1])
    # baseline parameters (:536-552)
    for r in range(n_mu):
        res = diff_M_term[r] - 2. * (sum_diff_f['mu'][0][r] / T_f)
        for i in range(d):
            res += 2. * (sum_diff_f['K_%d' % i][0][2][r] / T_f)
        grad[r] = res
    for ix_f, intervals in ((0, ker), (1, imp)):
        for p in range(d):
            for r in range(intervals[p][1] - intervals[p][0]):
                res = 0.
                for i in range(d):
                    if i != p:
                        res += 2. * (sum_diff_f['upsilon_%d_%d' % (i, p)][ix_f][2][r] / T_f)
                        res += 2. * (sum_diff_f['upsilon_%d_%d' % (p, i)][ix_f][1][r] / T_f)
                res += 2. * (sum_diff_f['upsilon_%d_%d' % (p, p)][ix_f][r] / T_f)
                res -= 2. * (sum_diff_f['phi_%d' % p][ix_f][r] / T_f)
                if ix_f == 0:
                    res += 2. * (sum_diff_f['K_%d' % p][0][1][r] / T_f)
                else:
                    res += 2. * (sum_diff_f['K_%d' % p][1][r] / T_f)
                res += sum_diff_f['upsilonzero_%d' % p][ix_f][r] / T_f
                grad[intervals[p][0] + r] = res
    return grad


def flatten(M_term, diff_M_term, sum_f, sum_diff_f, prefix):
    """Flat {key: array} view of an estimator state, for npz storage and comparisons."""
    out = {'%s/M_term' % prefix: np.float64(M_term),
           '%s/diff_M_term' % prefix: np.asarray(diff_M_term, dtype=float)}
    for name, val in sum_f.items():
        out['%s/f/%s' % (prefix, name)] = np.float64(val)
        for ix_f, part in enumerate(sum_diff_f[name]):
            if isinstance(part, list):
                for ix_func in (1, 2):
                    out['%s/df/%s/%d/%d' % (prefix, name, ix_f, ix_func)] = \
                        np.asarray(part[ix_func], dtype=float)
            else:
                out['%s/df/%s/%d' % (prefix, name, ix_f)] = np.asarray(part, dtype=float)
    return out
