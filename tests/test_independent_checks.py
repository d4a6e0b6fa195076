"""Checks of the LSE against algorithms that share nothing with the pairwise sums:
  * the reference's O(N) Markov recursion for exponential models
    (RecurrentExponential.get_lse_k, rec_exp.py:1692; golden values made by
    tests/golden/make_golden_recexp.py from the live reference) at ~10^4 events per type;
  * central finite differences of the LSE against the assembled gradient.
CPU: the numpy oracle (and the estimator's host assembly over it); GPU: the device."""
import types

import numpy as np
import pytest

from cases import CASES, case_x
from helpers import golden_times, load_golden
from oracle import exact as OE

RECEXP = ['recexp_exp1d', 'recexp_exp2d']


@pytest.mark.parametrize('name', RECEXP)
def test_oracle_lse_matches_the_recursion(name):
    g = load_golden(name)
    times, T_f = golden_times(g)
    model = CASES[str(g['case'])][0]
    for k in range(model['d']):
        x = g['k%d/x' % k]
        sum_f, _, _ = OE.exact_sums_windowed(model, k, x, times, T_f)
        lse = OE.recombine_lse_k(model, k, x, sum_f, T_f, len(times[k]) / T_f)
        assert abs(lse - g['k%d/lse' % k]) <= 1e-10 * abs(g['k%d/lse' % k])


def _estimator(engine, model, times, T_f, k):
    import aslsd_b200
    path = types.SimpleNamespace(list_times=times, T_f=T_f, d=len(times),
                                 n_events=np.array([len(t) for t in times]),
                                 eta=np.array([len(t) / T_f for t in times]), list_times2end=None,
                                 kappa=None, varpi=None, lag_sizes=None)
    est = aslsd_b200.MHPExactEstimB200(engine=engine, reference_recombine=False)
    est.initialize(k, 10, aslsd_b200.MHP(model), path)
    return est


def _finite_difference_check(engine, name, T_f, n_per_type):
    model = CASES[name][0]
    d = model['d']
    rng = np.random.default_rng(31)
    times = [np.sort(rng.uniform(0, T_f, size=n_per_type + 37 * i)) for i in range(d)]
    for k in range(d):
        est = _estimator(engine, model, times, T_f, k)
        x = np.asarray(case_x(name)[k], dtype=float)
        grad = est.lse_k_grad_estimate(x)
        for p in range(len(x)):
            h = 1e-5 * max(1., abs(x[p]))
            xp, xm = x.copy(), x.copy()
            xp[p] += h
            xm[p] -= h
            fd = (est.lse_k_estimate(xp) - est.lse_k_estimate(xm)) / (2. * h)
            assert abs(fd - grad[p]) <= 1e-5 * max(1., abs(grad[p])), (k, p, fd, grad[p])


@pytest.mark.parametrize('name', ['exp2d', 'exp1d2b'])
def test_gradient_matches_finite_differences_of_the_lse(name):
    """Exponential kernels (the reference's Gaussian derivative forms are transcribed as they
    are, mistakes included - DESIGN quirk 2 - and do not match finite differences upstream
    either); correct gradient assembly (reference_recombine=False)."""
    from oracle_engine import OracleEngine
    _finite_difference_check(OracleEngine(), name, 120., 150)


@pytest.mark.gpu
@pytest.mark.parametrize('name', RECEXP)
def test_device_lse_matches_the_recursion(name):
    import aslsd_b200
    g = load_golden(name)
    times, T_f = golden_times(g)
    model = CASES[str(g['case'])][0]
    engine = aslsd_b200.Engine(0)
    for k in range(model['d']):
        est = _estimator(engine, model, times, T_f, k)
        lse = est.lse_k_estimate(g['k%d/x' % k])
        # the LSE is a difference of sums ten times its size: 1e-10 per term -> 1e-9 here
        assert abs(lse - g['k%d/lse' % k]) <= 1e-9 * abs(g['k%d/lse' % k])
    engine.close()


@pytest.mark.gpu
def test_device_gradient_matches_finite_differences_of_the_lse():
    import aslsd_b200
    engine = aslsd_b200.Engine(0)
    _finite_difference_check(engine, 'exp2d', 600., 700)
    engine.close()
