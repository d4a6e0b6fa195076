"""Oracle for the per-event sums that follow a fit (test infrastructure): intensity at the jump
times (models/hawkes/linear/mhp.py:669-690) and residuals (stats/residual_analysis/
goodness_of_fit.py:134-236), restated with the reference's own O(N^2) loop structure on the plain
model description of oracle/exact.py."""
import numpy as np

from oracle import functionals as F


def intensity_at_jumps(model, mu, kernel_param, times):
    d = model['d']
    out = [float(mu[i]) + np.zeros(len(times[i])) for i in range(d)]
    for i in range(d):
        for j in range(d):
            ti, tj = times[i], times[j]
            for m in range(len(ti)):
                hi = m if i == j else np.searchsorted(tj, ti[m], side='left')
                if hi == 0:
                    continue
                out[i][m] += np.sum(F.phi(model['kernels'][i][j], ti[m] - tj[:hi],
                                          np.asarray(kernel_param[i][j], dtype=float)))
    return out


def residuals_k(model, k, mu, kernel_param, times):
    """goodness_of_fit.py:169-228: kernel_part[i][m] = sum of psi_ki over the type-i events
    strictly before t^k_m (0 where there is none), residual = Poisson part + differences."""
    d = model['d']
    tk = times[k]
    kernel_part = np.zeros((d, len(tk)))
    for i in range(d):
        ti = times[i]
        for m in range(len(tk)):
            hi = m if i == k else np.searchsorted(ti, tk[m], side='left')
            # quirk (:192-203): the value at varpi[k][i][1] is only computed when that index is
            # positive, so when the first type-k event already has type-i predecessors its
            # compensator value stays 0
            if hi > 0 and not (m == 0 and i != k):
                kernel_part[i][m] = np.sum(F.psi(model['kernels'][k][i], tk[m] - ti[:hi],
                                                 np.asarray(kernel_param[k][i], dtype=float)))
    poisson = float(mu[k]) * (tk[1:] - tk[:-1])
    return poisson + (kernel_part[:, 1:] - kernel_part[:, :-1]).sum(axis=0)


def residuals_k_mtlh(model, k, mu_k, kernel_param_k, expected_impact_k, times):
    """goodness_of_fit.py:245-301 for a constant baseline (compensator mu t,
    basis_baseline_constant.py): kernel_part[i][m] = e_ki x (sum of psi_ki over the type-i events
    strictly before t^k_m).  Unlike the MHP version above the value at the first type-k event that
    has type-i predecessors is always computed (:271-276), and the events carry no individual
    weight: the impact enters through its expectation e_ki only."""
    d = model['d']
    tk = times[k]
    kernel_part = np.zeros((d, len(tk)))
    for i in range(d):
        ti = times[i]
        for m in range(len(tk)):
            hi = m if i == k else np.searchsorted(ti, tk[m], side='left')
            if hi > 0:
                kernel_part[i][m] = expected_impact_k[i] * np.sum(
                    F.psi(model['kernels'][k][i], tk[m] - ti[:hi], np.asarray(kernel_param_k[i], dtype=float)))
    poisson = float(mu_k) * tk[1:] - float(mu_k) * tk[:-1]
    return poisson + (kernel_part[:, 1:] - kernel_part[:, :-1]).sum(axis=0)

