"""Interpolation error behind csrc/lag_quadrature.cuh: degree-(ND-1) Chebyshev interpolation of
the interior closed forms on bins of width h = fraction x (smallest scale of the row).

For every family of functions the lag quadrature integrates (Gaussian phi and its derivatives,
Gaussian upsilon value / cross and own-basis derivatives with the upper limit at +infinity,
exponential phi / upsilon) the script prints the largest interpolation error over all bins of the
lag window, relative to the largest value of the function, for a range of bin widths and node
counts.  The library uses ND = 12 and fraction <= 0.5 (LQ_ND, LQ_SCALE_FRACTION).

    python scripts/quadrature_error.py
"""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import ndtr


def pdf(x):
    return np.exp(-0.5 * x * x) / np.sqrt(2 * np.pi)


def worst(f, s_max, h, nd):
    i = np.arange(nd)
    x = np.cos(np.pi * (i + 0.5) / nd)
    xt = np.linspace(-1, 1, 401)
    err, top = 0., 0.
    for j in range(int(s_max / h) + 1):
        a = j * h
        nodes = a + h * (x + 1) / 2
        c = C.chebfit(x, f(nodes), nd - 1)
        st = a + h * (xt + 1) / 2
        ft = f(st)
        err = max(err, np.max(np.abs(C.chebval(xt, c) - ft)))
        top = max(top, np.max(np.abs(ft)))
    return err / top


def families(b1, d1, b2, d2):
    """(name, function of the lag, scale, window) for one Gaussian basis pair."""
    sig = np.hypot(b1, b2)
    bs = b1 * b2 / sig
    r1, r2 = b2 ** 2 / sig ** 2, b1 ** 2 / sig ** 2
    P = lambda s: pdf((s - (d2 - d1)) / sig) / sig
    lo = lambda s: -(r1 * d1 + r2 * (d2 - s)) / bs
    J = lambda s: 1 - ndtr(lo(s))
    scale = min(sig, b2 * sig / b1)
    s_max = max(d2 - d1, 0) + 9 * sig
    yield 'upsilon value', lambda s: P(s) * J(s), scale, s_max
    yield 'upsilon d/d beta1', lambda s: P(s) * ((b1 / sig ** 2) * (((s - d2 + d1) ** 2 / sig ** 2) - 1) * J(s)
                                                    + (b2 / (b1 * sig ** 2)) * (2 * b1 * (d2 - d1 - s) / sig ** 2
                                                                                + lo(s) * bs / b1) * pdf(lo(s))), scale, s_max
    yield 'upsilon d/d delta1', lambda s: P(s) * (((d2 - s - d1) / sig ** 2) * J(s)
                                                     + (b2 / (b1 * sig ** 2)) * pdf(lo(s))), scale, s_max
    yield 'phi value', lambda s: pdf((s - d1) / b1) / b1, b1, d1 + 9 * b1
    yield 'phi d/d beta', lambda s: pdf((s - d1) / b1) / b1 * (((s - d1) / b1) ** 2 - 1), b1, d1 + 9 * b1
    a = lambda s: -np.sqrt(2) * (d1 - 0.5 * s) / b1
    P0 = lambda s: pdf(s / (np.sqrt(2) * b1)) / (np.sqrt(2) * b1)
    yield 'own-basis d/d beta', lambda s: P0(s) * (((s * s - 2 * b1 * b1) / (2 * b1 ** 3)) * (1 - ndtr(a(s)))
                                                      + (np.sqrt(2) * (-d1 + 0.5 * s) / b1 ** 2) * pdf(a(s))), \
        np.sqrt(2) * b1, 9 * np.sqrt(2) * b1
    yield 'exp phi d/d beta', lambda s: np.exp(-s / b1) * (1 - s / b1), b1, 42 * b1


if __name__ == '__main__':
    cases = [(0.5, 0., 0.5, 9.), (0.5, 3., 0.5, 3.), (0.3, 1., 1.2, 2.5), (1.5, 0.5, 0.4, 4.)]
    print('%-22s' % 'family' + ''.join('  f=%.2f nd=%2d' % (f, nd) for f in (0.3, 0.5, 0.75) for nd in (10, 12, 14)))
    rows = {}
    for case in cases:
        for name, fun, scale, s_max in families(*case):
            errs = [worst(fun, s_max, f * scale, nd) for f in (0.3, 0.5, 0.75) for nd in (10, 12, 14)]
            rows[name] = np.maximum(rows.get(name, 0.), errs)
    for name, errs in rows.items():
        print('%-22s' % name + ''.join('      %8.1e' % e for e in errs))
