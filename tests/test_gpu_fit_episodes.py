"""MHP.fit_episodes (mhp.py:1806-2008) on the device: one resident context per path, iterates
equal to the same loop over the numpy oracle."""
import numpy as np
import pytest

from cases import CASES, case_x
from oracle_engine import OracleEngine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,exact', [('exp2d', True), ('gauss2d', True), ('exp2d', False)])
def test_fit_episodes_on_device_matches_oracle_loop(name, exact):
    import aslsd_b200
    from aslsd_b200.events import Episodes, ProcessPath
    model = CASES[name][0]
    d = model['d']
    rng = np.random.default_rng(5)
    T_f = 400.
    raw = [[np.sort(rng.uniform(0, T_f, size=500 + 40 * q + 30 * i)) for i in range(d)] for q in range(3)]
    x_0 = [np.asarray(v, float) for v in case_x(name)]
    kw = dict(n_iter=4, exact_grad=exact, seed=7)
    if not exact:
        kw.update(n_exact_single=50, n_samples_adaptive_single=60, nonadaptive_sample_size_single=80,
                  n_samples_adaptive_double=200, nonadaptive_sample_size_double=90)
    paths = [ProcessPath(t, T_f) for t in raw]
    mirror = aslsd_b200.MHP(model)
    x_dev, log_dev = mirror.fit_episodes(Episodes(paths, T_f=T_f), x_0, **kw)
    engines = {id(est.engine) for row in mirror.fit_estim for est in row}
    assert len(engines) == 3            # every path stays resident in its own context
    # the same loop over the oracle: per-path engines given explicitly
    cls = aslsd_b200.MHPExactEstimB200 if exact else aslsd_b200.MHPStratEstimB200
    o_paths = [ProcessPath(t, T_f, kappa=p.kappa, varpi=p.varpi, lag_sizes=p.lag_sizes, engine=OracleEngine())
               for t, p in zip(raw, paths)]
    ests = [[cls(engine=p.engine) for p in o_paths] for _ in range(d)]
    x_ora, log_ora = aslsd_b200.MHP(model).fit_episodes(Episodes(o_paths, T_f=T_f), x_0, estimators=ests,
                                                         **{k: v for k, v in kw.items() if k != 'exact_grad'})
    for k in range(d):
        np.testing.assert_allclose(log_dev['grad'][k], log_ora['grad'][k], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(x_dev[k], x_ora[k], rtol=1e-9)
