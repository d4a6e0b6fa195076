"""Oracle closed forms for the basis kernels on the LSE path (test infrastructure).

Plain numpy restatement of the reference's functionals, including the
derivative branches that are known to be mathematically wrong (SURVEY.md
section 4 quirk 10): parity means reproducing them.

Reference files followed:
  * ExponentialKernel     functionals/kernels/basis_kernels/basis_kernel_exponential.py
        phi/diff_phi :387-396, psi/diff_psi :402-411, upsilon :22-24,
        diff_cross_upsilon :55-83, diff_sim_upsilon :425-435
  * GaussianKernel        functionals/kernels/basis_kernels/basis_kernel_gaussian.py
        phi/diff_phi :448-467, psi/diff_psi :469-488, upsilon :21-34,
        diff_cross_upsilon :65-145, diff_sim_upsilon :502-525
  * leaf math             utilities/useful_functions.py
        gaussian_pdf :417, gaussian_int :509 (scipy.special.ndtr), share_ratio,
        prod_ratio :299-336
  * KernelModel sums      functionals/kernels/kernel.py:224-332
  * fixed variables       functionals/kernels/basis_kernel.py:364-401

A kernel is described by plain data: a list of bases, each
``{'type': 'exp' | 'gauss', 'fixed': {var_index: value}}``; its parameter
vector holds the non-fixed variables basis after basis, in variable order
(omega, beta[, delta]).
"""
import numpy as np
from scipy.special import ndtr

N_VARS = {'exp': 2, 'gauss': 3}
SQRT_2PI = np.sqrt(2 * np.pi)


# ----------------------------------------------------------------------------
# leaf math (useful_functions.py:417, :509)
# ----------------------------------------------------------------------------
def npdf(x, mean=0., std=1.):
    return 1. / (std * np.sqrt(2 * np.pi)) * np.exp(-0.5 * ((x - mean) / std) ** 2)


def nint(lo, hi):
    return ndtr(hi) - ndtr(lo)


# ----------------------------------------------------------------------------
# kernel description helpers
# ----------------------------------------------------------------------------
def basis_n_param(basis):
    return N_VARS[basis['type']] - len(basis.get('fixed', {}))


def kernel_n_param(kernel):
    return sum(basis_n_param(b) for b in kernel)


def basis_vars(basis, params):
    """Re-insert fixed variables (basis_kernel.py:364-401)."""
    fixed = basis.get('fixed', {})
    out = np.empty(N_VARS[basis['type']])
    it = iter(params)
    for v in range(len(out)):
        out[v] = fixed[v] if v in fixed else next(it)
    return out


def free_var_index(basis, ix_param):
    """Local parameter index -> variable index (BasisKernel.ix_map)."""
    fixed = basis.get('fixed', {})
    free = [v for v in range(N_VARS[basis['type']]) if v not in fixed]
    return free[ix_param]


def kernel_split(kernel, params):
    """Per-basis full variable vectors from a kernel parameter vector
    (kernel.py:97-117 interval_map + basis_kernel.py:364-401)."""
    out, pos = [], 0
    for b in kernel:
        n = basis_n_param(b)
        out.append(basis_vars(b, params[pos:pos + n]))
        pos += n
    return out


def kernel_locate(kernel, ix_param):
    """Kernel parameter index -> (basis index, variable index)
    (kernel.py:97-109 ix_map)."""
    pos = 0
    for r, b in enumerate(kernel):
        n = basis_n_param(b)
        if ix_param < pos + n:
            return r, free_var_index(b, ix_param - pos)
        pos += n
    raise IndexError(ix_param)


# ----------------------------------------------------------------------------
# ExponentialKernel
# ----------------------------------------------------------------------------
def exp_phi(t, v):
    return v[0] * v[1] * np.exp(-v[1] * t)


def exp_diff_phi(t, ix, v):
    if ix == 0:
        return v[1] * np.exp(-v[1] * t)
    if ix == 1:
        return v[0] * np.exp(-v[1] * t) * (1. - v[1] * t)
    raise ValueError(ix)


def exp_psi(t, v):
    return v[0] * (1 - np.exp(-v[1] * t))


def exp_diff_psi(t, ix, v):
    if ix == 0:
        return 1 - np.exp(-v[1] * t)
    if ix == 1:
        return v[0] * t * np.exp(-v[1] * t)
    raise ValueError(ix)


def exp_exp_upsilon(t, s, v1, v2):
    w1, b1 = v1
    w2, b2 = v2
    return (w1 * w2 * ((b1 * b2) / (b1 + b2)) * np.exp(-b2 * s)
            * (1 - np.exp(-(b1 + b2) * t)))


def exp_exp_diff_cross(t, s, ix_func, ix, v1, v2):
    w1, b1 = v1
    w2, b2 = v2
    bsum = b1 + b2
    if ix_func == 1:
        if ix == 0:
            return (w2 * ((b1 * b2) / bsum) * np.exp(-b2 * s)
                    * (1. - np.exp(-bsum * t)))
        if ix == 1:
            return (w1 * w2 * np.exp(-b2 * s)
                    * ((b2 / (b2 + b1)) ** 2 * (1. - np.exp(-(b2 + b1) * t))
                       + (b2 * b1) / (b2 + b1) * t * np.exp(-(b2 + b1) * t)))
    if ix_func == 2:
        if ix == 0:
            return (w1 * ((b1 * b2) / bsum) * np.exp(-b2 * s)
                    * (1. - np.exp(-bsum * t)))
        if ix == 1:
            return (w1 * w2 * np.exp(-b2 * s)
                    * ((b1 / (b1 + b2)) ** 2 * (1. - np.exp(-bsum * t))
                       + (b2 * b1) / (b2 + b1)
                       * ((t + s) * np.exp(-bsum * t) - s)))
    raise ValueError((ix_func, ix))


def exp_diff_sim(t, s, ix, v):
    w, b = v
    if ix == 0:
        return w * b * np.exp(-b * s) * (1 - np.exp(-2 * b * t))
    if ix == 1:
        return (0.5 * w ** 2 * np.exp(-b * s)
                * (1. - b * s + np.exp(-2 * b * t) * (b * (2 * t + s) - 1)))
    raise ValueError(ix)


# ----------------------------------------------------------------------------
# GaussianKernel
# ----------------------------------------------------------------------------
def gauss_phi(t, v):
    return v[0] * npdf(t, v[2], v[1])


def gauss_diff_phi(t, ix, v):
    w, b, dl = v
    if ix == 0:
        return npdf(t, dl, b)
    if ix == 1:
        return (w / b) * npdf(t, dl, b) * (((t - dl) / b) ** 2 - 1.)
    if ix == 2:
        return w * ((t - dl) / b ** 2) * npdf(t, dl, b)
    raise ValueError(ix)


def gauss_psi(t, v):
    w, b, dl = v
    return w * nint(-dl / b, (t - dl) / b)


def gauss_diff_psi(t, ix, v):
    w, b, dl = v
    if ix == 0:
        return nint(-dl / b, (t - dl) / b)
    if ix == 1:
        return (w / b ** 2) * ((dl - t) * npdf((t - dl) / b)
                               - dl * npdf(-dl / b))
    if ix == 2:
        return -(w / b) * (npdf((t - dl) / b) - npdf(-dl / b))
    raise ValueError(ix)


def _gg_common(s, v1, v2):
    _, b1, d1 = v1
    _, b2, d2 = v2
    sig = np.sqrt(b1 ** 2 + b2 ** 2)
    b_s = (b1 * b2) / sig
    d_s = ((b2 ** 2 / (b2 ** 2 + b1 ** 2)) * d1
           + (b1 ** 2 / (b1 ** 2 + b2 ** 2)) * (d2 - s))
    return sig, b_s, d_s


def gauss_gauss_upsilon(t, s, v1, v2):
    sig, b_s, d_s = _gg_common(s, v1, v2)
    return (v1[0] * v2[0] * npdf(s, v2[2] - v1[2], sig)
            * nint(-d_s / b_s, (t - d_s) / b_s))


def gauss_gauss_diff_cross(t, s, ix_func, ix, v1, v2):
    w1, b1, d1 = v1
    w2, b2, d2 = v2
    sig, b_s, d_s = _gg_common(s, v1, v2)
    P = npdf(s, d2 - d1, sig)
    J = nint(-d_s / b_s, (t - d_s) / b_s)
    p_hi = npdf((t - d_s) / b_s)
    p_lo = npdf(-d_s / b_s)
    if ix_func == 1:
        if ix == 0:
            return w2 * P * J
        if ix == 1:
            cdf_term = (b1 / sig ** 2) * (((s - d2 + d1) ** 2 / sig ** 2) - 1) * J
            c = 2 * b1 * (d2 - d1 - s) / sig ** 2
            pdf_term = -((b2 / (b1 * sig ** 2))
                         * ((c + ((t - d_s) / b1)) * p_hi
                            - (c - d_s / b1) * p_lo))
            return w1 * w2 * P * (cdf_term + pdf_term)
        if ix == 2:
            cdf_term = ((d2 - s - d1) / sig ** 2) * J
            pdf_term = -((b2 / (b1 * sig ** 2)) * (p_hi - p_lo))
            return w1 * w2 * P * (cdf_term + pdf_term)
    if ix_func == 2:
        if ix == 0:
            return w1 * P * J
        if ix == 1:
            cdf_term = (b2 / sig ** 2) * (((s - d2 + d1) ** 2 / sig ** 2) - 1) * J
            c = 2 * b2 * (d1 - d2 + s) / sig ** 2
            pdf_term = -((b1 / (b2 * sig ** 2))
                         * ((c - ((t - d_s) / b2)) * p_hi
                            - (c + d_s / b2) * p_lo))
            return w2 * w1 * P * (cdf_term + pdf_term)
        if ix == 2:
            cdf_term = ((d1 + s - d2) / sig ** 2) * J
            pdf_term = -((b1 / (b2 * sig ** 2)) * (p_hi - p_lo))
            return w2 * w1 * P * (cdf_term + pdf_term)
    raise ValueError((ix_func, ix))


def gauss_diff_sim(t, s, ix, v):
    w, b, dl = v
    r2 = np.sqrt(2.)
    P0 = npdf(s, 0., r2 * b)
    if ix == 0:
        # verbatim: no s/2 shift in the limits (basis_kernel_gaussian.py:507-508)
        return 2. * w * P0 * nint(-r2 * dl / b, r2 * (t - dl) / b)
    lo = -r2 * (dl - 0.5 * s) / b
    hi = r2 * (t - dl + 0.5 * s) / b
    if ix == 1:
        cdf_term = ((s ** 2 - 2 * b ** 2) / (2 * b ** 3)) * nint(lo, hi)
        pdf_term = -((r2 * (t - dl + 0.5 * s) / b ** 2) * npdf(hi)
                     - (r2 * (-dl + 0.5 * s) / b ** 2) * npdf(lo))
        return w ** 2 * P0 * (cdf_term + pdf_term)
    if ix == 2:
        return -((w ** 2 * r2 / b) * P0 * (npdf(hi) - npdf(lo)))
    raise ValueError(ix)


# ----------------------------------------------------------------------------
# dispatch by basis type (the reference dispatches on str(type(.)),
# basis_kernel_exponential.py:418-420)
# ----------------------------------------------------------------------------
_PHI = {'exp': exp_phi, 'gauss': gauss_phi}
_DPHI = {'exp': exp_diff_phi, 'gauss': gauss_diff_phi}
_PSI = {'exp': exp_psi, 'gauss': gauss_psi}
_DPSI = {'exp': exp_diff_psi, 'gauss': gauss_diff_psi}
# ----------------------------------------------------------------------------
# mixed families: verbatim, including the overflow guards
# ----------------------------------------------------------------------------
def exp_gauss_upsilon(t, s, v1, v2):
    """basis_kernel_exponential.py:30-39"""
    w1, b1 = v1[0], v1[1]
    w2, b2, d2 = v2
    d_s = d2 - s - b1 * b2 ** 2
    return (w1 * w2 * b1 * np.exp(-b1 * (d2 - s - 0.5 * b1 * b2 ** 2))
            * nint(-d_s / b2, (t - d_s) / b2))


def exp_gauss_diff_cross(t, s, ix_func, ix, v1, v2):
    """basis_kernel_exponential.py:89-156"""
    w1, b1 = v1[0], v1[1]
    w2, b2, d2 = v2
    d_s = d2 - s - b1 * b2 ** 2
    arg = -b1 * (d2 - s - 0.5 * b1 * b2 ** 2)
    with np.errstate(over='ignore', invalid='ignore'):
        e = np.exp(arg)
        J = nint(-d_s / b2, (t - d_s) / b2)
        if ix_func == 1:
            if ix == 0:
                return np.where(arg > 10 ** 2, 0., w2 * b1 * e * J)
            if ix == 1:
                cdf_term = (1. / b1 - d2 + s + 0.5 * b1 * b2 ** 2) * J
                pdf_term = b2 * (npdf((t - d_s) / b2) - npdf(-d_s / b2))
                return np.where(arg > 10 ** 2, 0., w1 * w2 * b1 * e * (cdf_term + pdf_term))
        elif ix_func == 2:
            if ix == 0:
                return np.where(arg > 10 ** 2, 0., w1 * b1 * e * J)
            if ix == 1:
                cdf_term = (b2 * b1 ** 2) * J
                pdf_term = ((b1 + (d2 - s - t) / b2 ** 2) * npdf((t - d_s) / b2)
                            - (b1 + (d2 - s) / b2 ** 2) * npdf(-d_s / b2))
                return np.where(arg > 10 ** 2, 0., w2 * w1 * b1 * e * (cdf_term + pdf_term))
            if ix == 2:
                cdf_term = b1 * J
                pdf_term = (1. / b2) * (npdf((t - d_s) / b2) - npdf(-d_s / b2))
                return -np.where(arg > 10 ** 2, 0., w2 * w1 * b1 * e * (cdf_term + pdf_term))
    raise ValueError((ix_func, ix))


def gauss_exp_upsilon(t, s, v1, v2):
    """basis_kernel_gaussian.py:40-50"""
    w1, b1, d1 = v1
    w2, b2 = v2[0], v2[1]
    d_s = d1 - b2 * b1 ** 2
    return (w1 * w2 * b2 * np.exp(-b2 * (d1 + s - 0.5 * b2 * b1 ** 2))
            * nint(-d_s / b1, (t - d_s) / b1))


def gauss_exp_diff_cross(t, s, ix_func, ix, v1, v2):
    """basis_kernel_gaussian.py:151-213 (the beta_1 / delta_1 branches shift d_s by s:
    transcribed as they are)"""
    w1, b1, d1 = v1
    w2, b2 = v2[0], v2[1]
    arg = -b2 * (d1 + s - 0.5 * b2 * b1 ** 2)
    e = np.exp(arg)
    if ix_func == 1:
        if ix == 0:
            d_s = d1 - b2 * b1 ** 2
            return np.where(arg > 10 ** 2, 0., w2 * b2 * e * nint(-d_s / b1, (t - d_s) / b1))
        if ix == 1:
            d_s = d1 + s - b2 * b1 ** 2
            cdf_term = b1 * (b2 ** 2) * nint(-d_s / b1, (t - d_s) / b1)
            pdf_term = ((b2 - ((t - d1) / b1 ** 2)) * npdf((t - d_s) / b1)
                        - (b2 + (d1 / b1 ** 2)) * npdf(-d_s / b1))
            return w1 * w2 * b2 * e * (cdf_term + pdf_term)
        if ix == 2:
            d_s = (d1 + s - b2 * b1 ** 2) / b1
            cdf_term = -b2 * nint(-d_s, t / b1 - d_s)
            pdf_term = (-1. / b1) * (npdf(t / b1 - d_s) - npdf(-d_s))
            return w1 * w2 * b2 * e * (cdf_term + pdf_term)
    elif ix_func == 2:
        if ix == 0:
            d_s = d1 - b2 * b1 ** 2
            return w1 * b2 * e * nint(-d_s / b1, (t - d_s) / b1)
        if ix == 1:
            d_s = (d1 - b2 * b1 ** 2) / b1
            cdf_term = ((1. / b2) - d1 - s + 0.5 * b2 * b1 ** 2) * nint(-d_s, (t / b1) - d_s)
            pdf_term = b1 * (npdf((t / b1) - d_s) - npdf(-d_s))
            return w2 * w1 * b2 * e * (cdf_term + pdf_term)
    raise ValueError((ix_func, ix))


_UPS = {('exp', 'exp'): exp_exp_upsilon, ('gauss', 'gauss'): gauss_gauss_upsilon,
        ('exp', 'gauss'): exp_gauss_upsilon, ('gauss', 'exp'): gauss_exp_upsilon}
_DCROSS = {('exp', 'exp'): exp_exp_diff_cross,
           ('gauss', 'gauss'): gauss_gauss_diff_cross,
           ('exp', 'gauss'): exp_gauss_diff_cross,
           ('gauss', 'exp'): gauss_exp_diff_cross}
_DSIM = {'exp': exp_diff_sim, 'gauss': gauss_diff_sim}


def _pair(b1, b2):
    key = (b1['type'], b2['type'])
    if key not in _UPS:
        raise NotImplementedError("basis pair %s x %s is outside the oracle's "
                                  "scope (SURVEY section 8f rank 3)" % key)
    return key


# ----------------------------------------------------------------------------
# KernelModel level (kernel.py:224-332)
# ----------------------------------------------------------------------------
def phi(kernel, t, params):
    res = 0.
    for b, v in zip(kernel, kernel_split(kernel, params)):
        res = res + _PHI[b['type']](t, v)
    return res


def diff_phi(kernel, t, ix_param, params):
    r, iv = kernel_locate(kernel, ix_param)
    return _DPHI[kernel[r]['type']](t, iv, kernel_split(kernel, params)[r])


def psi(kernel, t, params):
    res = 0.
    for b, v in zip(kernel, kernel_split(kernel, params)):
        res = res + _PSI[b['type']](t, v)
    return res


def diff_psi(kernel, t, ix_param, params):
    r, iv = kernel_locate(kernel, ix_param)
    return _DPSI[kernel[r]['type']](t, iv, kernel_split(kernel, params)[r])


def upsilon(kernel_1, kernel_2, t, s, params_1, params_2):
    """Sum over all basis pairs (kernel.py:257-272)."""
    V1 = kernel_split(kernel_1, params_1)
    V2 = kernel_split(kernel_2, params_2)
    res = 0.
    for b1, v1 in zip(kernel_1, V1):
        for b2, v2 in zip(kernel_2, V2):
            res = res + _UPS[_pair(b1, b2)](t, s, v1, v2)
    return res


def diff_sim_upsilon(kernel, t, s, ix_param, params):
    """Own-basis closed form + cross terms with every other basis in both slots
    (kernel.py:274-300)."""
    V = kernel_split(kernel, params)
    r, iv = kernel_locate(kernel, ix_param)
    res = _DSIM[kernel[r]['type']](t, s, iv, V[r])
    for q in range(len(kernel)):
        if q != r:
            res = res + _DCROSS[_pair(kernel[r], kernel[q])](t, s, 1, iv, V[r], V[q])
    for p in range(len(kernel)):
        if p != r:
            res = res + _DCROSS[_pair(kernel[p], kernel[r])](t, s, 2, iv, V[p], V[r])
    return res


def diff_cross_upsilon(kernel_1, kernel_2, t, s, ix_func, ix_param,
                       params_1, params_2):
    """kernel.py:302-332."""
    V1 = kernel_split(kernel_1, params_1)
    V2 = kernel_split(kernel_2, params_2)
    res = 0.
    if ix_func == 1:
        r, iv = kernel_locate(kernel_1, ix_param)
        for q in range(len(kernel_2)):
            res = res + _DCROSS[_pair(kernel_1[r], kernel_2[q])](
                t, s, 1, iv, V1[r], V2[q])
        return res
    if ix_func == 2:
        r, iv = kernel_locate(kernel_2, ix_param)
        for p in range(len(kernel_1)):
            res = res + _DCROSS[_pair(kernel_1[p], kernel_2[r])](
                t, s, 2, iv, V1[p], V2[r])
        return res
    raise ValueError(ix_func)
