"""Saturated samples of the single-index terms of the stratified estimator.

Far from T_f the per-sample value of ``psi_i`` and of ``upsilonzero_i`` is one constant: every
upper integration limit has reached +infinity in fp64 (``ndtr(z) == 1.0`` for z >= 9,
``1 - exp(-x) == 1.0`` for x >= 40).  Most single-index adaptive strata of a long path lie entirely
in that regime; the reference's allocation statistic there (``np.std`` of the function values,
optimize/estimators/mhp_stratified_estimator.py:631-635) is the rounding residue of numpy's
summation of that constant, and the integer sample counts of the next iteration are a truncation
of it (:587, :637-645).  A constant that differs in its last bit - another ``ndtr``, another
association of a sum over basis kernels - changes the residue, hence the draws of every later
iteration.

So the constant is not recomputed on the device: it is taken from the model's own closures
(``model.psi[k][i]``, ``model.upsilon[i][i][k]``, the reference's Python, evaluated once per term and
call at a large t with the argument shapes the reference uses) and handed to the sampled-lag
kernel with the threshold beyond which it applies (``aslsd_sample_term.sat_t / v_sat``).  Below the
threshold the device's closed forms are used: those strata have a genuine variance.

The mirror model (``aslsd_b200.model.MHP``, used where the reference is absent) carries the same
closures through ``value_closures`` below: ``psi`` and ``upsilon(t, 0)`` for exponential and
Gaussian basis kernels, composed over the bases as ``KernelModel`` does
(functionals/kernels/kernel.py:241-267; basis_kernel_exponential.py:22-24, :402-403;
basis_kernel_gaussian.py:21-34, :469-473; utilities/useful_functions.py:417, :509, share_ratio).
"""
import itertools

import numpy as np

Z_SAT = 9.0      # ndtr(z) == 1.0 exactly for z >= 8.3 (scipy and CUDA); margin
X_SAT = 40.0     # 1 - exp(-x) == 1.0 exactly for x >= 37


def _basis_vars(bases, params):
    """Full variables of every basis of a kernel (BasisKernel.make_vars, basis_kernel.py:364-401)."""
    out, pos = [], 0
    for type_name, free_mask, fixed in bases:
        nv = 2 if type_name == 'ExponentialKernel' else 3
        v = []
        for q in range(nv):
            if (free_mask >> q) & 1:
                v.append(float(params[pos]))
                pos += 1
            else:
                v.append(float(fixed[q]))
        out.append((type_name, v))
    return out


def psi_threshold(bases, params):
    """Time to the end beyond which psi of the kernel is its limit in fp64 (None: unknown family)."""
    t = 0.
    for type_name, v in _basis_vars(bases, params):
        if type_name == 'ExponentialKernel':
            t = max(t, X_SAT / v[1])
        elif type_name == 'GaussianKernel':
            t = max(t, v[2] + Z_SAT * v[1])
        else:
            return None
    return t


def upsilonzero_threshold(bases, params):
    """The same for upsilon(t, 0) of the kernel with itself (all basis pairs)."""
    t = 0.
    V = _basis_vars(bases, params)
    for (ty1, v1), (ty2, v2) in itertools.product(V, V):
        if ty1 == 'ExponentialKernel' and ty2 == 'ExponentialKernel':
            t = max(t, X_SAT / (v1[1] + v2[1]))
        elif ty1 == 'GaussianKernel' and ty2 == 'GaussianKernel':
            var = v1[1] ** 2 + v2[1] ** 2
            b_s = v1[1] * v2[1] / np.sqrt(var)
            d_s = (v2[1] ** 2 / var) * v1[2] + (v1[1] ** 2 / var) * v2[2]
            t = max(t, d_s + Z_SAT * b_s)
        else:
            return None      # mixed pairs: no saturation constant, the device's values stay
    return t


def term_saturation(model, desc, k, i, kind, x_k):
    """(sat_t, v_sat) of the single-index term (kind 0: psi_i, 1: upsilonzero_i) of sub-problem k at
    x_k, or None when the model has no closures / an unknown family."""
    d = desc.d
    lo, hi = desc.interval_map[k][i]
    x_ki = np.asarray(x_k[lo:hi], dtype=np.float64)
    if kind == 0:
        closure = getattr(model, 'psi', None)
        thr = psi_threshold(desc.bases[k][i], x_ki)
    else:
        closure = getattr(model, 'upsilon', None)
        # the value closure of every upsilon is bound to the structure of K[d-1][d-1]
        # (mhp.py:527-536, late binding)
        thr = upsilonzero_threshold(desc.bases[d - 1][d - 1], x_ki)
    if closure is None or thr is None:
        return None
    t_far = np.array([1e6 * max(thr, 1.)])
    try:
        if kind == 0:
            v = float(x_k[0]) * closure[k][i](t_far, x_ki)          # K = param_mu * psi (:1106-1108)
        else:
            v = closure[i][i][k](t_far, 0., x_ki, x_ki)             # :1148-1157
    except Exception:
        return None
    v = np.asarray(v, dtype=np.float64).ravel()
    if v.size != 1 or not np.isfinite(v[0]):
        return None
    return float(thr), float(v[0])


# ---- value closures of the mirror model ------------------------------------------------------------
def _gaussian_pdf(t, mean=0., std=1.):
    return 1. / (std * np.sqrt(2 * np.pi)) * np.exp(-0.5 * ((t - mean) / std) ** 2)


def _gaussian_int(start, end):
    from scipy.special import ndtr
    return ndtr(end) - ndtr(start)


def _share_ratio(x, y):
    return x / (x + y)


def _basis_psi(type_name, t, v):
    if type_name == 'ExponentialKernel':
        return v[0] * (1 - np.exp(-v[1] * t))
    return v[0] * _gaussian_int(-v[2] / v[1], (t - v[2]) / v[1])


def _basis_upsilon(ty1, ty2, t, s, v1, v2):
    if ty1 == 'ExponentialKernel' and ty2 == 'ExponentialKernel':
        return (v1[0] * v2[0] * ((v1[1] * v2[1]) / (v1[1] + v2[1]))
                * np.exp(-v2[1] * s) * (1 - np.exp(-(v1[1] + v2[1]) * t)))
    if ty1 == 'GaussianKernel' and ty2 == 'GaussianKernel':
        sum_beta = np.sqrt(v1[1] ** 2 + v2[1] ** 2)
        b_s = (v1[1] * v2[1]) / sum_beta
        d_s = (_share_ratio(v2[1] ** 2, v1[1] ** 2) * v1[2]
               + _share_ratio(v1[1] ** 2, v2[1] ** 2) * (v2[2] - s))
        return (v1[0] * v2[0] * _gaussian_pdf(s, v2[2] - v1[2], sum_beta)
                * _gaussian_int(-d_s / b_s, (t - d_s) / b_s))
    raise NotImplementedError('mixed basis pairs have no saturation constant')


def value_closures(desc):
    """(psi, upsilon) in the reference's layout: psi[k][i](t, params), upsilon[i][j][k](t, s,
    params_1, params_2) - the latter with the structure of K[d-1][d-1] for every (i, j, k)."""
    d = desc.d

    def make_psi(bases):
        def psi(t, params):
            res = 0.
            for type_name, v in _basis_vars(bases, params):
                res += _basis_psi(type_name, t, v)
            return res
        return psi

    last = desc.bases[d - 1][d - 1]

    def upsilon(t, s, params_1, params_2):
        res = 0.
        for (ty1, v1), (ty2, v2) in itertools.product(_basis_vars(last, params_1), _basis_vars(last, params_2)):
            res += _basis_upsilon(ty1, ty2, t, s, v1, v2)
        return res
    psi = [[make_psi(desc.bases[k][i]) for i in range(d)] for k in range(d)]
    ups = [[[upsilon for _ in range(d)] for _ in range(d)] for _ in range(d)]
    return psi, ups
