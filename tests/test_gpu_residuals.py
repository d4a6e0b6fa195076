"""Intensity at the jumps and residuals on the device (aslsd_point_sums) against the golden
vectors of the reference and, at a larger size, against the oracle.  Tolerance 1e-10."""
import numpy as np
import pytest

from cases import CASES
from helpers import RTOL, load_golden
from oracle import residuals as OR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
    import aslsd_b200
    return aslsd_b200.Engine(0)


@pytest.mark.parametrize('name', ['exp2d', 'gauss1d', 'sbf2d3', 'gauss2d'])
def test_matches_reference_golden(engine, name):
    import aslsd_b200
    from aslsd_b200 import residuals as R
    model, mu, kp, T_f = CASES[name]
    g = load_golden('residuals_' + name)
    d = model['d']
    times = [g['times/%d' % i] for i in range(d)]
    path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
    intensity = R.get_intensity_at_jumps(model, path, mu, kp, engine=engine)
    residuals = R.get_residuals(model, path, mu, kp, engine=engine)
    for i in range(d):
        np.testing.assert_allclose(intensity[i], g['intensity/%d' % i], rtol=RTOL, atol=1e-13)
        # a residual is a difference of compensator values: tolerance relative to those
        scale = np.max(np.abs(g['residuals/%d' % i]))
        np.testing.assert_allclose(residuals[i], g['residuals/%d' % i], rtol=RTOL, atol=1e-10 * scale)


def test_matches_oracle_on_a_longer_path(engine):
    """Long enough for the window truncation (phi) and the saturation shortcut (psi) to act:
    most earlier events are outside the kernels' support."""
    import aslsd_b200
    from aslsd_b200 import residuals as R
    name = 'sbf2d3'
    model, mu, kp, _ = CASES[name]
    rng = np.random.default_rng(6)
    T_f = 3000.
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (2500, 1800)]
    path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
    intensity = R.get_intensity_at_jumps(model, path, mu, kp, engine=engine)
    want = OR.intensity_at_jumps(model, mu, kp, times)
    for i in range(2):
        np.testing.assert_allclose(intensity[i], want[i], rtol=RTOL, atol=1e-13)
        got = R.get_residuals_k(model, path, i, mu, kp, engine=engine)
        ref = OR.residuals_k(model, i, mu, kp, times)
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)
        # time-change theorem sanity: residuals of a Poisson-like path under this model are O(1)
        assert 0.2 < np.mean(got) < 5.
