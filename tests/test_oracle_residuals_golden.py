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
