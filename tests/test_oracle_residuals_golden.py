"""The residual / intensity oracle (oracle/residuals.py) against golden vectors generated from
the reference (tests/golden/make_golden_residuals.py)."""
import numpy as np
import pytest

from cases import CASES
from helpers import load_golden
from oracle import residuals as OR

NAMES = ['exp2d', 'gauss1d', 'sbf2d3', 'gauss2d']


@pytest.mark.parametrize('name', NAMES)
def test_oracle_matches_reference(name):
    model, mu, kp, T_f = CASES[name]
    g = load_golden('residuals_' + name)
    d = model['d']
    times = [g['times/%d' % i] for i in range(d)]
    intensity = OR.intensity_at_jumps(model, mu, kp, times)
    for i in range(d):
        np.testing.assert_allclose(intensity[i], g['intensity/%d' % i], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(OR.residuals_k(model, i, mu, kp, times), g['residuals/%d' % i],
                                   rtol=1e-10, atol=1e-12)


MTLH_NAMES = ['mtlh_exp2d', 'mtlh_gauss1d', 'mtlh_sbf2d3']


def _mtlh_golden(name):
    from cases import MTLH_CASES
    model = MTLH_CASES[name][0]
    g = load_golden('residuals_' + name)
    d = model['d']
    times = [g['times/%d' % i] for i in range(d)]
    kp = [[g['kernel_param/%d/%d' % (k, i)] for i in range(d)] for k in range(d)]
    mu = [g['mu/%d' % k] for k in range(d)]
    return model, g, d, times, mu, kp, g['expected_impact']


@pytest.mark.parametrize('name', MTLH_NAMES)
def test_oracle_matches_reference_marked_models(name):
    """goodness_of_fit.get_residuals_mtlh: expected impacts, no first-residual quirk."""
    model, g, d, times, mu, kp, E = _mtlh_golden(name)
    for k in range(d):
        got = OR.residuals_k_mtlh(model, k, mu[k][0], kp[k], E[k], times)
        np.testing.assert_allclose(got, g['residuals/%d' % k], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize('name', MTLH_NAMES)
def test_mirror_matches_reference_marked_models(name):
    """The host side of aslsd_b200.get_residuals_mtlh over the oracle's point sums."""
    import types
    import aslsd_b200
    from oracle_engine import OracleEngine
    model, g, d, times, mu, kp, E = _mtlh_golden(name)
    path = types.SimpleNamespace(list_times=times, T_f=float(g['T_f']))
    got = aslsd_b200.get_residuals_mtlh({'d': d, 'kernels': model['kernels']}, path, mu, kp, E,
                                        engine=OracleEngine())
    for k in range(d):
        np.testing.assert_allclose(got[k], g['residuals/%d' % k], rtol=1e-10, atol=1e-12)
