"""Device-side cluster simulation (csrc/simulate.cu; the generation loop of MHP.simulate,
models/hawkes/linear/mhp.py:733-854).  The law is the reference's, the random numbers are not, so
the checks are distributional: stationary rates eta = (I - A)^-1 mu, the lag law of the offspring
seen through the fitted LSE gradient at the truth, agreement with the numpy simulator of
aslsd_b200/simulate.py, determinism in the seed."""
import numpy as np
import pytest
from scipy.special import ndtr

from cases import CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
    import aslsd_b200
    return aslsd_b200.Engine(0)


def _rates(model, mu, kp):
    from aslsd_b200.model import describe_model
    desc = describe_model(model)
    d = desc.d
    A = np.zeros((d, d))
    for i in range(d):
        for j in range(d):
            pos = 0
            for type_name, free_mask, fixed in desc.bases[i][j]:
                nv = 2 if type_name == 'ExponentialKernel' else 3
                vals = []
                for q in range(nv):
                    if (free_mask >> q) & 1:
                        vals.append(kp[i][j][pos])
                        pos += 1
                    else:
                        vals.append(fixed[q])
                # children with a non-positive lag are dropped (mhp.py:826-838): a Gaussian basis
                # keeps the share Phi(delta / beta) of its mass
                A[i, j] += vals[0] * (1. if nv == 2 else ndtr(vals[2] / vals[1]))
    return np.linalg.solve(np.eye(d) - A, np.asarray(mu, float))


@pytest.mark.parametrize('name', ['exp1d', 'exp2d', 'gauss1d', 'sbf2d3', 'mix2d', 'exp3d'])
def test_simulated_paths_have_the_models_law(engine, name):
    from aslsd_b200 import simulate
    model, mu, kp, _ = CASES[name]
    d = model['d']
    T_f = 40000.
    times = engine.simulate(model, mu, kp, T_f, seed=7)
    assert len(times) == d
    eta = _rates(model, mu, kp)
    for i, t in enumerate(times):
        assert np.all(np.diff(t) >= 0.) and t[0] >= 0. and t[-1] < T_f
        # counts: mean eta T_f; the variance of a Hawkes count exceeds Poisson's by 1 / (1 - ||A||)^2
        assert abs(len(t) - eta[i] * T_f) <= 8. * np.sqrt(eta[i] * T_f) / (1. - 0.8), (name, i, len(t), eta[i] * T_f)
    # same seed, same path; another seed, another path
    again = engine.simulate(model, mu, kp, T_f, seed=7)
    other = engine.simulate(model, mu, kp, T_f, seed=8)
    for a, b, c in zip(times, again, other):
        np.testing.assert_array_equal(a, b)
        assert len(a) != len(c) or np.any(a != c)
    # against the numpy simulator of the same law: inter-arrival distributions.  The gaps of a
    # clustered process are dependent, so the two-sample KS p-value is not calibrated: the
    # statistic itself is bounded (sampling noise ~ 1.4 sqrt(2 / n) x the cluster size <= 5)
    from scipy.stats import ks_2samp
    host = simulate.simulate_mhp(model, mu, kp, T_f, seed=11)
    for a, b in zip(times, host):
        stat = ks_2samp(np.diff(a), np.diff(b)).statistic
        assert stat < 10. * np.sqrt(2. / min(len(a), len(b))), (name, stat, len(a), len(b))
        assert abs(len(a) - len(b)) <= 10. * np.sqrt(len(b)) / (1. - 0.8), (name, len(a), len(b))


def test_lse_gradient_vanishes_at_the_truth(engine):
    """The exact LSE gradient of a long simulated path is small at the generating parameters and
    not one standard step away: the offspring lags follow the kernels' shapes."""
    import aslsd_b200
    name = 'exp2d'
    model, mu, kp, _ = CASES[name]
    T_f = 2.0e5
    times = engine.simulate(model, mu, kp, T_f, seed=3)
    path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
    mirror = aslsd_b200.MHP(model)
    for k in range(2):
        x_true = np.concatenate([[mu[k]]] + [np.asarray(kp[k][i], float) for i in range(2)])
        est = aslsd_b200.MHPExactEstimB200(engine=engine, reference_recombine=False)
        est.initialize(k, 2, mirror, path)
        g0 = est.lse_k_grad_estimate(x_true)
        x_off = x_true * np.array([1.2] + [0.8, 1.25] * 2)
        g1 = est.lse_k_grad_estimate(x_off)
        assert np.linalg.norm(g0) < 0.1 * np.linalg.norm(g1), (g0, g1)


def test_simulation_rejects_unstable_models(engine):
    import aslsd_b200
    model = CASES['exp1d'][0]
    with pytest.raises(aslsd_b200.EngineError):
        engine.simulate(model, [1.0], [[[1.2, 1.0]]], 1000.)     # branching ratio 1.2
    with pytest.raises(aslsd_b200.EngineError):
        engine.simulate(model, [1.0], [[[0.5, -1.0]]], 1000.)


def test_benchmark_size_stream(engine):
    """~10^7 events of the C5 model: sorted, right rates."""
    from aslsd_b200 import simulate
    model = simulate.sbf_gauss_model(10, 10)
    w = simulate.random_mixture_weights(10, seed=1234)
    kp = [[0.5 * w / 10 for _ in range(10)] for _ in range(10)]
    T_f = 1.1e6
    times = simulate.simulate_mhp_device(model, np.full(10, 0.5), kp, T_f, seed=1234, engine=engine)
    n = np.array([len(t) for t in times])
    # half of the children of the delta = 0 basis fall before their parent and are dropped:
    # the numpy simulator of the same law gives 9 966 445 events (bench.py's stream)
    assert abs(n.sum() - 9.97e6) < 5e4 and np.all(np.abs(n - 9.97e5) < 1e4), n
    for t in times:
        assert np.all(np.diff(t) >= 0.)
