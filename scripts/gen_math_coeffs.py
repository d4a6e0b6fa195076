"""Generate the polynomial coefficients of csrc/fast_math.cuh (run once; needs mpmath).

exp:   Chebyshev interpolant (near-minimax) of e^r on |r| <= ln(2)/2, degree 12.
erfc:  (1 + 2a) exp(a^2) erfc(a) as a polynomial of degree 22 in t = (a - 4)/(a + 4),
       a >= 0 (t in [-1, 1)), Chebyshev interpolant converted to the monomial basis
       in 60-digit arithmetic.
"""
import mpmath as mp

mp.mp.dps = 60


def cheb_monomial(fun, deg):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    fv = [fun(x) for x in nodes]
    c = [sum(fv[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n)) * 2 / n
         for j in range(n)]
    c[0] /= 2
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        new = [mp.mpf(0)] + [2 * x for x in T[j - 1]]
        for i, x in enumerate(T[j - 2]):
            new[i] -= x
        T.append(new)
    mono = [mp.mpf(0)] * n
    for j in range(n):
        for i, x in enumerate(T[j]):
            mono[i] += c[j] * x
    return mono


def emit(name, coefs):
    print('__constant__ double %s[%d] = {' % (name, len(coefs)))
    for c in coefs:
        print('    %s,' % mp.nstr(c, 20))
    print('};')


half = mp.log(2) / 2
exp_c = cheb_monomial(lambda u: mp.exp(u * half), 12)
exp_c = [c / half ** i for i, c in enumerate(exp_c)]     # back to the variable r = u * half
emit('FM_EXP', exp_c)


def g(t):
    if t >= 1:
        return 2 / mp.sqrt(mp.pi)
    a = 4 * (1 + t) / (1 - t)
    return (1 + 2 * a) * mp.exp(a * a) * mp.erfc(a)


emit('FM_ERFC', cheb_monomial(g, 22))
