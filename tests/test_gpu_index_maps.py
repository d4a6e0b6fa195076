"""Bit-exact parity of the batched searchsorted kernel (kappa / varpi /
lag_sizes) against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

from helpers import load_golden, golden_times
from oracle import time_ordering as OT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
    import aslsd_b200
    return aslsd_b200.Engine(0)


def _assert_maps(engine, times, T_f, kappa, varpi, lags):
    d = len(times)
    engine.set_events(times, T_f)
    g_kappa, g_varpi, g_lags = engine.index_maps()
    for i in range(d):
        for j in range(d):
            assert g_kappa[i][j].dtype == np.int64
            np.testing.assert_array_equal(g_kappa[i][j], kappa[i][j], err_msg='kappa %d %d' % (i, j))
            np.testing.assert_array_equal(g_varpi[i][j], varpi[i][j], err_msg='varpi %d %d' % (i, j))
            np.testing.assert_array_equal(g_lags[i][j], lags[i][j], err_msg='lags %d %d' % (i, j))


def test_reference_golden_index_maps(engine):
    g = load_golden('index_maps')
    for name in g['names']:
        d = int(g['%s/d' % name])
        times = [g['%s/times/%d' % (name, i)] for i in range(d)]
        kappa = [[g['%s/kappa/%d/%d' % (name, i, j)] for j in range(d)] for i in range(d)]
        varpi = [[g['%s/varpi/%d/%d' % (name, i, j)] for j in range(d)] for i in range(d)]
        lags = [[g['%s/lag_sizes/%d/%d' % (name, i, j)] for j in range(d)] for i in range(d)]
        _assert_maps(engine, times, float(g['%s/T_f' % name]), kappa, varpi, lags)


@pytest.mark.parametrize('name', ['exp2d', 'gauss2d', 'sbf2d3', 'exp1d'])
def test_simulated_path_index_maps(engine, name):
    g = load_golden('exact_' + name)
    times, T_f = golden_times(g)
    d = len(times)
    kappa = [[g['kappa/%d/%d' % (i, j)] for j in range(d)] for i in range(d)]
    varpi = [[g['varpi/%d/%d' % (i, j)] for j in range(d)] for i in range(d)]
    lags = [[g['lag_sizes/%d/%d' % (i, j)] for j in range(d)] for i in range(d)]
    _assert_maps(engine, times, T_f, kappa, varpi, lags)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_random_tie_heavy_paths_against_oracle(engine, seed):
    rng = np.random.default_rng(seed)
    d = int(rng.integers(2, 6))
    times = []
    for i in range(d):
        n = int(rng.integers(1, 3000))
        if seed == 0:   # quantised times: many ties across and within types
            t = np.sort(rng.integers(0, 2000, size=n)).astype(float) * 0.125
        else:
            t = np.sort(rng.uniform(0, 500, size=n))
        times.append(t)
    T_f = 500.
    kappa, varpi = OT.kappa_varpi_closed_form(times, T_f)
    lags = OT.lag_sizes(times, kappa, varpi)
    _assert_maps(engine, times, T_f, kappa, varpi, lags)


def test_large_path_properties(engine):
    """10^6 events per type: closed form vs numpy and the reference's property
    checkers on a sample of positions (time_ordering_test.py:23-66)."""
    rng = np.random.default_rng(7)
    times = [np.sort(rng.uniform(0, 1e6, size=n)) for n in (1_000_000, 700_000)]
    T_f = 1e6
    engine.set_events(times, T_f)
    kappa, varpi, lags = engine.index_maps()
    o_kappa, o_varpi = OT.kappa_varpi_closed_form(times, T_f)
    for i in range(2):
        for j in range(2):
            np.testing.assert_array_equal(kappa[i][j], o_kappa[i][j])
            np.testing.assert_array_equal(varpi[i][j], o_varpi[i][j])
    # property: t^j[kappa] < t^i_m <= t^j[kappa+1]
    m = rng.integers(0, len(times[0]), size=2000)
    q = kappa[1][0][m]
    ok = q >= 0
    assert np.all(times[1][q[ok]] < times[0][m[ok]])
    nxt = np.minimum(q + 1, len(times[1]) - 1)
    assert np.all((times[0][m] <= times[1][nxt]) | (q + 1 >= len(times[1])))
    np.testing.assert_array_equal(lags[0][1][1:], len(times[0]) - varpi[0][1][1:])


def test_process_path_mirror(engine):
    import aslsd_b200
    g = load_golden('exact_exp2d')
    times, T_f = golden_times(g)
    path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
    assert path.d == 2
    np.testing.assert_array_equal(path.n_events, [len(t) for t in times])
    np.testing.assert_array_equal(path.kappa[0][1], g['kappa/0/1'])
    np.testing.assert_array_equal(path.list_times2end[1], T_f - times[1])
    with pytest.raises(IndexError):
        aslsd_b200.ProcessPath([times[0], np.array([])], T_f, engine=engine)
