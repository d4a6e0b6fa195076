"""Synthetic event streams for benchmarks and large parity inputs.

Cluster (branching) simulation of a linear Hawkes process, generation by
generation and fully vectorised in numpy: immigrants of type i are a Poisson
process of rate mu_i; every event of type j spawns Poisson(omega_ijr) children
of type i per basis r of phi_ij, at lags drawn from the basis' normalised
shape (exponential(beta) or N(delta, beta) restricted to positive lags).

This plays the role of ``MHP.simulate`` (models/hawkes/linear/mhp.py:733-854)
as the source of inputs; it is not bit-compatible with it (the golden vectors
of tests/golden use the reference's own simulator) and is not part of the
accelerated path.
"""
import numpy as np

from aslsd_b200.model import describe_model


def simulate_mhp(model, mu, kernel_param, T_f, seed=1234, max_generations=200):
    """Returns ``list_times`` (sorted float64 arrays, one per event type).

    kernel_param[i][j] is the parameter vector of phi_ij (free variables only,
    as in the reference)."""
    desc = describe_model(model)
    d = desc.d
    rng = np.random.default_rng(seed)
    # full variables of every basis
    basis_vars = [[[] for _ in range(d)] for _ in range(d)]
    for i in range(d):
        for j in range(d):
            params = iter(np.asarray(kernel_param[i][j], dtype=float))
            for type_name, free_mask, fixed in desc.bases[i][j]:
                nv = 2 if type_name == 'ExponentialKernel' else 3
                v = [next(params) if (free_mask >> q) & 1 else fixed[q] for q in range(nv)]
                basis_vars[i][j].append((type_name, v))
    events = [[] for _ in range(d)]
    current = [np.sort(rng.uniform(0., T_f, size=rng.poisson(mu[i] * T_f))) for i in range(d)]
    for _ in range(max_generations):
        if sum(len(c) for c in current) == 0:
            break
        for i in range(d):
            events[i].append(current[i])
        children = [[] for _ in range(d)]
        for j in range(d):
            parents = current[j]
            if len(parents) == 0:
                continue
            for i in range(d):
                for type_name, v in basis_vars[i][j]:
                    n_child = rng.poisson(v[0], size=len(parents))
                    tot = int(n_child.sum())
                    if tot == 0:
                        continue
                    origin = np.repeat(parents, n_child)
                    if type_name == 'ExponentialKernel':
                        lag = rng.exponential(1. / v[1], size=tot)
                    else:
                        lag = rng.normal(v[2], v[1], size=tot)
                    t = origin + lag
                    keep = (lag > 0.) & (t < T_f)
                    children[i].append(t[keep])
        current = [np.concatenate(c) if c else np.empty(0) for c in children]
    out = []
    for i in range(d):
        t = np.sort(np.concatenate(events[i])) if events[i] else np.empty(0)
        # strictly increasing times within a type (orderly process)
        if len(t) > 1:
            dup = np.flatnonzero(np.diff(t) <= 0.)
            while len(dup):
                t[dup + 1] = np.nextafter(t[dup + 1], np.inf)
                t = np.sort(t)
                dup = np.flatnonzero(np.diff(t) <= 0.)
        out.append(t)
    return out


def simulate_mhp_device(model, mu, kernel_param, T_f, seed=1234, engine=None):
    """The same law simulated on the device (``Engine.simulate``, csrc/simulate.cu): 10^7 events
    in tens of milliseconds instead of the seconds of the numpy generations above."""
    if engine is None:
        from aslsd_b200.events import default_engine
        engine = default_engine()
    return engine.simulate(model, mu, kernel_param, T_f, seed=seed)


def sbf_gauss_model(d, n_basis=10, beta=0.5, spacing=1.0):
    """d-dimensional MHP whose kernels are sums of n_basis Gaussians with
    (beta, delta) fixed and omega free (fit_sbgauss1d10r.py:60-63)."""
    kernel = [{'type': 'gauss', 'fixed': {1: beta, 2: spacing * r}} for r in range(n_basis)]
    return {'d': d, 'kernels': [[[dict(b) for b in kernel] for _ in range(d)] for _ in range(d)]}


def random_mixture_weights(n, coeff=1., seed=1234):
    """get_random_mixture_weights of utilities/useful_statistics.py:23-32."""
    rng = np.random.default_rng(seed)
    w = rng.uniform(size=n)
    w /= np.sum(w)
    w *= coeff
    w[-1] = coeff - np.sum(w[:-1])
    return w


def config_c5(d=10, n_basis=10, T_f=1e6, seed=1234):
    """Config C5 of SURVEY.md section 8d: d-dimensional SBF-Gaussian MHP, every
    adjacency row sums to 0.5, mu_i = 0.5 (eta_i ~ 1)."""
    model = sbf_gauss_model(d, n_basis)
    w = random_mixture_weights(n_basis, seed=1234)
    kernel_param = [[0.5 * w / d for _ in range(d)] for _ in range(d)]
    mu = np.full(d, 0.5)
    times = simulate_mhp(model, mu, kernel_param, T_f, seed=seed)
    rng = np.random.default_rng(12)
    xs = []
    for k in range(d):
        x = np.concatenate([[mu[k]]] + [kernel_param[k][i] for i in range(d)])
        xs.append(x * (1 + 0.1 * rng.uniform(-1, 1, size=len(x))))
    return model, times, T_f, xs
