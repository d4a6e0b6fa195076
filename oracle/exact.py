"""Oracle for the exact LSE sums of an MHP sub-problem k (test infrastructure).

Restates ``MHPExactEstim.compute_objective`` and the five ``compute_*_sum``
methods of the reference (optimize/estimators/mhp_exact_estimator.py:543-761)
and the LSE / gradient assembly (``recombine_*``; the stratified estimator's
loop mhp_stratified_estimator.py:976-1068 is followed, because the exact
estimator's own gradient loop :518-530 overwrites entries for d >= 2, SURVEY
quirk 1).

Two evaluation strategies:
  * ``exact_sums_literal``  - a Python loop over target events m with numpy over
    *all* earlier sources n, exactly like the reference (O(N^2), small inputs).
  * ``exact_sums_windowed`` - only pairs whose lag s = t_m - t_n lies within the
    kernels' effective support are evaluated (agrees with
    the literal sum to ~1e-15 relative); used at sizes where O(N^2) is not
    feasible and as the CPU baseline of bench.py.

Model description (plain data): ``model = {'d': d, 'kernels': K}`` with
``K[k][i]`` the kernel phi_ki as a list of bases (see oracle/functionals.py).
``x_k = [mu_k, params(phi_k0), ..., params(phi_k,d-1)]``.

Late binding (models/hawkes/linear/mhp.py:527-536): every
``upsilon[i][j][k]`` / ``diff_cross_upsilon[i][j][k]`` closure of the reference
ends up bound to the kernel objects of the last loop iteration, K[d-1][d-1];
``ups_kernel()`` reproduces that.  ``diff_sim_upsilon[k][i]`` is bound to
K[k][i] itself (mhp.py:513-514).
"""
import numpy as np

from oracle import functionals as F

Z_GAUSS = 9.0      # Gaussian support half-width in standard deviations
C_EXP = 42.0       # exponential support in units of 1/beta


def interval_map(model, k):
    """[left, right) of phi_ki's parameters in x_k (mhp.py:276-299)."""
    out, pos = [], 1
    for i in range(model['d']):
        n = F.kernel_n_param(model['kernels'][k][i])
        out.append((pos, pos + n))
        pos += n
    return out


def n_param_k(model, k):
    return interval_map(model, k)[-1][1]


def ups_kernel(model):
    d = model['d']
    return model['kernels'][d - 1][d - 1]


def _null_sums(model, k):
    d = model['d']
    K = model['kernels']
    sum_f, sum_diff_f = {}, {}
    for i in range(d):
        n_i = F.kernel_n_param(K[k][i])
        for name in ('phi_%d' % i, 'upsilonzero_%d' % i):
            sum_f[name] = 0.
            sum_diff_f[name] = np.zeros(n_i)
        sum_f['psi_%d' % i] = 0.
        sum_diff_f['psi_%d' % i] = [None, np.zeros(n_i), np.zeros(1)]
    for i in range(d):
        for j in range(d):
            name = 'upsilon_%d_%d' % (i, j)
            sum_f[name] = 0.
            if i == j:
                sum_diff_f[name] = np.zeros(F.kernel_n_param(K[k][i]))
            else:
                sum_diff_f[name] = [None, np.zeros(F.kernel_n_param(K[k][i])),
                                    np.zeros(F.kernel_n_param(K[k][j]))]
    return sum_f, sum_diff_f


# ----------------------------------------------------------------------------
# single sums (mhp_exact_estimator.py:543-609)
# ----------------------------------------------------------------------------
def _single_sums(model, k, x_k, times, T_f, sum_f, sum_diff_f, ranges=None):
    d = model['d']
    K = model['kernels']
    imap = interval_map(model, k)
    mu_k = x_k[0]
    KU = ups_kernel(model)
    for i in range(d):
        x_ki = x_k[imap[i][0]:imap[i][1]]
        lo, hi = (0, len(times[i])) if ranges is None else ranges[i]
        t = T_f - times[i][lo:hi]
        n_i = F.kernel_n_param(K[k][i])
        # psi (:543-578)
        name = 'psi_%d' % i
        s_psi = np.sum(F.psi(K[k][i], t, x_ki))
        sum_f[name] = mu_k * s_psi
        for p in range(n_i):
            sum_diff_f[name][1][p] = mu_k * np.sum(F.diff_psi(K[k][i], t, p, x_ki))
        sum_diff_f[name][2][0] = s_psi
        # upsilon at lag zero (:580-609)
        name = 'upsilonzero_%d' % i
        sum_f[name] = np.sum(F.upsilon(KU, KU, t, 0., x_ki, x_ki))
        for p in range(n_i):
            sum_diff_f[name][p] = np.sum(
                F.diff_sim_upsilon(K[k][i], t, 0., p, x_ki))


# ----------------------------------------------------------------------------
# literal double sums (mhp_exact_estimator.py:611-707)
# ----------------------------------------------------------------------------
def exact_sums_literal(model, k, x_k, times, T_f):
    """All per-term sums with the reference's own loop structure."""
    x_k = np.asarray(x_k, dtype=float)
    d = model['d']
    K = model['kernels']
    KU = ups_kernel(model)
    imap = interval_map(model, k)
    sum_f, sum_diff_f = _null_sums(model, k)
    _single_sums(model, k, x_k, times, T_f, sum_f, sum_diff_f)
    for i in range(d):
        x_ki = x_k[imap[i][0]:imap[i][1]]
        n_i = F.kernel_n_param(K[k][i])
        # phi term (:611-638): targets of type k, sources of type i
        name = 'phi_%d' % i
        tk, ti = times[k], times[i]
        for m in range(len(tk)):
            hi = m if i == k else np.searchsorted(ti, tk[m], side='left')
            if hi == 0:
                continue
            s = tk[m] - ti[:hi]
            sum_f[name] += np.sum(F.phi(K[k][i], s, x_ki))
            for p in range(n_i):
                sum_diff_f[name][p] += np.sum(F.diff_phi(K[k][i], s, p, x_ki))
        # same-type upsilon (:675-707)
        name = 'upsilon_%d_%d' % (i, i)
        for m in range(1, len(ti)):
            s = ti[m] - ti[:m]
            t = (T_f - ti[m]) * np.ones(m)
            sum_f[name] += np.sum(F.upsilon(KU, KU, t, s, x_ki, x_ki))
            for p in range(n_i):
                sum_diff_f[name][p] += np.sum(
                    F.diff_sim_upsilon(K[k][i], t, s, p, x_ki))
    # cross upsilon (:640-673)
    for i in range(d):
        for j in range(d):
            if i == j:
                continue
            x_ki = x_k[imap[i][0]:imap[i][1]]
            x_kj = x_k[imap[j][0]:imap[j][1]]
            name = 'upsilon_%d_%d' % (i, j)
            ti, tj = times[i], times[j]
            for m in range(len(ti)):
                hi = np.searchsorted(tj, ti[m], side='left')
                if hi == 0:
                    continue
                s = ti[m] - tj[:hi]
                t = (T_f - ti[m]) * np.ones(hi)
                sum_f[name] += np.sum(F.upsilon(KU, KU, t, s, x_ki, x_kj))
                for ix_func, kern in ((1, K[k][i]), (2, K[k][j])):
                    for p in range(F.kernel_n_param(kern)):
                        sum_diff_f[name][ix_func][p] += np.sum(
                            F.diff_cross_upsilon(KU, KU, t, s, ix_func, p,
                                                 x_ki, x_kj))
    return sum_f, sum_diff_f


# ----------------------------------------------------------------------------
# windowed double sums
# ----------------------------------------------------------------------------
def basis_phi_support(b, v):
    """Lags outside [lo, hi] contribute < 1e-17 of the basis' peak."""
    if b['type'] == 'exp':
        return 0., C_EXP / v[1]
    return max(0., v[2] - Z_GAUSS * v[1]), v[2] + Z_GAUSS * v[1]


def basis_ups_support(b1, v1, b2, v2):
    if b2['type'] == 'exp':
        # exp x exp and gauss x exp decay like exp(-beta_2 s)
        return 0., C_EXP / v2[1]
    if b1['type'] == 'exp':
        # exp x gauss: the Gaussian factor sits at u + s around delta_2 with u >= 0
        return 0., max(0., v2[2]) + Z_GAUSS * v2[1]
    sig = np.sqrt(v1[1] ** 2 + v2[1] ** 2)
    dd = v2[2] - v1[2]
    return max(0., dd - Z_GAUSS * sig), dd + Z_GAUSS * sig


def phi_window(kernel, params):
    V = F.kernel_split(kernel, params)
    return max(basis_phi_support(b, v)[1] for b, v in zip(kernel, V))


def ups_window(kernel_1, kernel_2, params_1, params_2):
    V1 = F.kernel_split(kernel_1, params_1)
    V2 = F.kernel_split(kernel_2, params_2)
    w = 0.
    for b1, v1 in zip(kernel_1, V1):
        for b2, v2 in zip(kernel_2, V2):
            w = max(w, basis_ups_support(b1, v1, b2, v2)[1])
        # own-basis derivative closed forms use pdf(s; 0, sqrt(2) beta)
        w = max(w, basis_ups_support(b1, v1, b1, v1)[1])
    return w


def window_pairs(t_tgt, t_src, window, same, m_lo=0, m_hi=None,
                 chunk=1 << 16):
    """Yield (m, n) index arrays of all pairs with t_src[n] < t_tgt[m]
    (n < m when ``same``) and t_tgt[m] - t_src[n] <= window, m in [m_lo, m_hi)."""
    if m_hi is None:
        m_hi = len(t_tgt)
    for a in range(m_lo, m_hi, chunk):
        b = min(a + chunk, m_hi)
        tm = t_tgt[a:b]
        hi = (np.arange(a, b) if same
              else np.searchsorted(t_src, tm, side='left'))
        lo = np.minimum(np.searchsorted(t_src, tm - window, side='left'), hi)
        cnt = hi - lo
        tot = int(cnt.sum())
        if tot == 0:
            continue
        m_ix = np.repeat(np.arange(a, b), cnt)
        start = np.repeat(lo - (np.cumsum(cnt) - cnt), cnt)
        n_ix = np.arange(tot) + start
        yield m_ix, n_ix


def count_window_pairs(t_tgt, t_src, window, same, m_lo=0, m_hi=None):
    if m_hi is None:
        m_hi = len(t_tgt)
    tm = t_tgt[m_lo:m_hi]
    hi = (np.arange(m_lo, m_hi) if same
          else np.searchsorted(t_src, tm, side='left'))
    lo = np.minimum(np.searchsorted(t_src, tm - window, side='left'), hi)
    return int((hi - lo).sum())


def exact_sums_windowed(model, k, x_k, times, T_f, ranges=None,
                        count_only=False, chunk=1 << 15):
    """Same sums with lag truncation; ``ranges[i] = (m_lo, m_hi)`` restricts
    the *target* events of every term to a block (time-block shard).
    Returns (sum_f, sum_diff_f, n_pair_terms)."""
    x_k = np.asarray(x_k, dtype=float)
    d = model['d']
    K = model['kernels']
    KU = ups_kernel(model)
    imap = interval_map(model, k)
    sum_f, sum_diff_f = _null_sums(model, k)
    n_pairs = {}
    if not count_only:
        _single_sums(model, k, x_k, times, T_f, sum_f, sum_diff_f, ranges)

    def rng_of(a):
        return (0, len(times[a])) if ranges is None else ranges[a]

    for i in range(d):
        x_ki = x_k[imap[i][0]:imap[i][1]]
        n_i = F.kernel_n_param(K[k][i])
        name = 'phi_%d' % i
        W = phi_window(K[k][i], x_ki)
        lo, hi = rng_of(k)
        n_pairs[name] = count_window_pairs(times[k], times[i], W, i == k, lo, hi)
        if not count_only:
            for m_ix, n_ix in window_pairs(times[k], times[i], W, i == k,
                                           lo, hi, chunk):
                s = times[k][m_ix] - times[i][n_ix]
                sum_f[name] += np.sum(F.phi(K[k][i], s, x_ki))
                for p in range(n_i):
                    sum_diff_f[name][p] += np.sum(F.diff_phi(K[k][i], s, p, x_ki))
    for i in range(d):
        for j in range(d):
            x_ki = x_k[imap[i][0]:imap[i][1]]
            x_kj = x_k[imap[j][0]:imap[j][1]]
            name = 'upsilon_%d_%d' % (i, j)
            W = ups_window(KU, KU, x_ki, x_kj)
            lo, hi = rng_of(i)
            n_pairs[name] = count_window_pairs(times[i], times[j], W, i == j,
                                               lo, hi)
            if count_only:
                continue
            for m_ix, n_ix in window_pairs(times[i], times[j], W, i == j,
                                           lo, hi, chunk):
                s = times[i][m_ix] - times[j][n_ix]
                t = T_f - times[i][m_ix]
                sum_f[name] += np.sum(F.upsilon(KU, KU, t, s, x_ki, x_kj))
                if i == j:
                    for p in range(F.kernel_n_param(K[k][i])):
                        sum_diff_f[name][p] += np.sum(
                            F.diff_sim_upsilon(K[k][i], t, s, p, x_ki))
                else:
                    for ix_func, kern in ((1, K[k][i]), (2, K[k][j])):
                        for p in range(F.kernel_n_param(kern)):
                            sum_diff_f[name][ix_func][p] += np.sum(
                                F.diff_cross_upsilon(KU, KU, t, s, ix_func, p,
                                                     x_ki, x_kj))
    return sum_f, sum_diff_f, n_pairs


# ----------------------------------------------------------------------------
# assembly (mhp_stratified_estimator.py:976-1068; mu term :1073-1079)
# ----------------------------------------------------------------------------
def recombine_lse_k(model, k, x_k, sum_f, T_f, eta_k):
    d = model['d']
    mu_k = x_k[0]
    res = mu_k ** 2 - 2. * eta_k * mu_k
    for i in range(d):
        res += 2. * (sum_f['psi_%d' % i] / T_f)
    for i in range(d):
        for j in range(d):
            res += 2. * (sum_f['upsilon_%d_%d' % (i, j)] / T_f)
    for j in range(d):
        res -= 2. * (sum_f['phi_%d' % j] / T_f)
    for i in range(d):
        res += sum_f['upsilonzero_%d' % i] / T_f
    return res


def recombine_grad_lse_k(model, k, x_k, sum_diff_f, T_f, eta_k):
    d = model['d']
    imap = interval_map(model, k)
    grad = np.zeros(n_param_k(model, k))
    g0 = 2. * (x_k[0] - eta_k)
    for i in range(d):
        g0 += 2. * (sum_diff_f['psi_%d' % i][2][0] / T_f)
    grad[0] = g0
    for p in range(d):
        for r in range(imap[p][1] - imap[p][0]):
            res = 0.
            for i in range(d):
                if i != p:
                    res += 2. * (sum_diff_f['upsilon_%d_%d' % (i, p)][2][r] / T_f)
            for j in range(d):
                if j != p:
                    res += 2. * (sum_diff_f['upsilon_%d_%d' % (p, j)][1][r] / T_f)
            res += 2. * (sum_diff_f['upsilon_%d_%d' % (p, p)][r] / T_f)
            res -= 2. * (sum_diff_f['phi_%d' % p][r] / T_f)
            res += 2. * (sum_diff_f['psi_%d' % p][1][r] / T_f)
            res += sum_diff_f['upsilonzero_%d' % p][r] / T_f
            grad[imap[p][0] + r] = res
    return grad
