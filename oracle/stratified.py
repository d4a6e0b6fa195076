"""Oracle for the stratified (ASLSD) estimates (test infrastructure).

Restates, for one term and one stratum, the bodies of
``MHPStratEstim.estimate_sum_{adaptive,nonadaptive,exact}dom``
(optimize/estimators/mhp_stratified_estimator.py:534-900): function and
derivative values on the sampled indices (index -> time transforms of the
``compute_*_sum`` closures, :1097-1100, :1148-1151, :1196-1203, :1245-1252,
:1300-1307), their per-stratum sums / means, the allocation statistic (f, or
the 2-norm of the gradient when ``grad_alloc``; :605-635) and the allocation
update (:637-645).  Values come from oracle/functionals.py.
"""
import numpy as np

from oracle import exact as OE
from oracle import functionals as F

KIND_PSI, KIND_UPSILONZERO, KIND_PHI, KIND_UPSILON = 0, 1, 2, 3


def term_values(model, k, x_k, times, T_f, kind, i, j, ix_m, ix_n=None):
    """Per-sample outputs, shape (n_samples, n_out): column 0 is f, then the
    derivatives with respect to the parameters of ix_func 1, then ix_func 2
    (psi: [mu_k psi, mu_k d psi / d theta_ki ..., psi])."""
    x_k = np.asarray(x_k, dtype=float)
    K = model['kernels']
    KU = OE.ups_kernel(model)
    imap = OE.interval_map(model, k)
    x_ki = x_k[imap[i][0]:imap[i][1]]
    ix_m = np.asarray(ix_m, dtype=np.int64)
    cols = []
    if kind == KIND_PSI:
        t = T_f - times[i][ix_m]
        psi = F.psi(K[k][i], t, x_ki) + 0. * t
        cols.append(x_k[0] * psi)
        for p in range(F.kernel_n_param(K[k][i])):
            cols.append(x_k[0] * F.diff_psi(K[k][i], t, p, x_ki) + 0. * t)
        cols.append(psi)
    elif kind == KIND_UPSILONZERO:
        t = T_f - times[i][ix_m]
        cols.append(F.upsilon(KU, KU, t, 0., x_ki, x_ki) + 0. * t)
        for p in range(F.kernel_n_param(K[k][i])):
            cols.append(F.diff_sim_upsilon(K[k][i], t, 0., p, x_ki) + 0. * t)
    elif kind == KIND_PHI:
        s = times[k][ix_m] - times[i][np.asarray(ix_n, dtype=np.int64)]
        cols.append(F.phi(K[k][i], s, x_ki) + 0. * s)
        for p in range(F.kernel_n_param(K[k][i])):
            cols.append(F.diff_phi(K[k][i], s, p, x_ki) + 0. * s)
    elif kind == KIND_UPSILON:
        x_kj = x_k[imap[j][0]:imap[j][1]]
        t = T_f - times[i][ix_m]
        s = times[i][ix_m] - times[j][np.asarray(ix_n, dtype=np.int64)]
        cols.append(F.upsilon(KU, KU, t, s, x_ki, x_kj) + 0. * s)
        if i == j:
            for p in range(F.kernel_n_param(K[k][i])):
                cols.append(F.diff_sim_upsilon(K[k][i], t, s, p, x_ki) + 0. * s)
        else:
            for ix_func, kern in ((1, K[k][i]), (2, K[k][j])):
                for p in range(F.kernel_n_param(kern)):
                    cols.append(F.diff_cross_upsilon(KU, KU, t, s, ix_func, p, x_ki, x_kj) + 0. * s)
    else:
        raise ValueError(kind)
    return np.stack(cols, axis=1)


def stratum_statistics(vals, grad_alloc):
    """(sums, stat_mean, stat_std) of one stratum (:594-635): np.sum per output,
    and mean / population std of f or of ||grad f||_2."""
    if grad_alloc:
        stat = np.sqrt(np.sum(vals[:, 1:] ** 2, axis=1))
    else:
        stat = vals[:, 0]
    return vals.sum(axis=0), float(np.mean(stat)), float(np.std(stat))


def update_allocation(strata_sizes, stat_std, default_alloc):
    """rel_alloc of :637-645."""
    sizes = np.asarray(strata_sizes, dtype=float)
    with np.errstate(divide='ignore', invalid='ignore'):
        rel = sizes * np.sqrt(sizes / (sizes - 1.)) * np.asarray(stat_std, dtype=float)
    tot = np.sum(rel)
    return rel / tot if tot > 0. else np.asarray(default_alloc, dtype=float).copy()


def numpy_pairwise_sum(a):
    """numpy's float64 add.reduce restated in scalar Python (pairwise summation:
    blocks of at most 128 elements with 8 partial sums, halves split at a
    multiple of 8).  The CUDA stat / allocation kernels restate the same
    recursion; tests check this function bit for bit against np.sum."""
    a = np.asarray(a, dtype=np.float64)
    n = len(a)
    if n < 8:
        res = np.float64(0.)
        for v in a:
            res = res + v
        return res
    if n <= 128:
        r = [a[q] for q in range(8)]
        i = 8
        while i < n - (n % 8):
            for q in range(8):
                r[q] = r[q] + a[i + q]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:])
