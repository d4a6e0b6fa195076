"""s/epoch of an ASLSD fit on the GPU box (BASELINE.json metric M2; SURVEY 8d: an epoch is one
lse_k_grad_estimate + solver step for every k), for configs C2-C4, next to the exact-gradient
iteration (weight cache on).  Prints one JSON line per config."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import aslsd_b200
from aslsd_b200 import simulate

exp = lambda: [{'type': 'exp'}]
CONFIGS = {
    # C2: 2-D exponential (fit_exp2d.py:34-51), ~10^5 events
    'C2': ({'d': 2, 'kernels': [[exp(), exp()], [exp(), exp()]]}, [1.5, 1.],
           [[[.2, 1.], [.6, 1.5]], [[.7, 2.], [.1, 1.3]]], 7.7e3),
    # C3: 1-D Gaussian (fit_gauss1d.py:31-34), ~10^6 events
    'C3': ({'d': 1, 'kernels': [[[{'type': 'gauss'}]]]}, [1.5], [[[.5, .5, 3.]]], 3.33e5),
    # C4: 1-D SBF-10 model on a 3-Gaussian truth scaled to ~10^6 events
    'C4': (simulate.sbf_gauss_model(1, 10), [1.0],
           [[simulate.random_mixture_weights(10, coeff=0.5)]], 5e5),
}


def main():
    names = sys.argv[1:] or list(CONFIGS)
    engine = aslsd_b200.Engine(0)
    for name in names:
        model, mu, kp, T_f = CONFIGS[name]
        d = model['d']
        times = simulate.simulate_mhp(model, mu, kp, T_f, seed=1234)
        t0 = time.perf_counter()
        path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
        t_maps = time.perf_counter() - t0
        mirror = aslsd_b200.MHP(model)
        rng = np.random.default_rng(12)
        x0 = []
        for k in range(d):
            x = np.concatenate([[mu[k]]] + [np.asarray(kp[k][i], float) for i in range(d)])
            x0.append(x * (1 + 0.1 * rng.uniform(-1, 1, size=len(x))))
        out = {'config': name, 'd': d, 'n_events': int(sum(len(t) for t in times)),
               'index_maps_s': t_maps}
        for label, exact in (('aslsd', False), ('exact', True)):
            n_iter = 30 if not exact else 10
            mirror.fit(path, [x.copy() for x in x0], n_iter=3, exact_grad=exact, engine=engine, seed=1)
            t0 = time.perf_counter()
            mirror.fit(path, [x.copy() for x in x0], n_iter=n_iter, exact_grad=exact, engine=engine, seed=1)
            out['%s_s_per_epoch' % label] = (time.perf_counter() - t0) / n_iter
        print(json.dumps(out))


if __name__ == '__main__':
    main()
