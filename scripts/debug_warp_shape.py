import sys
sys.path.insert(0, '.')
import numpy as np
import aslsd_b200
from aslsd_b200.engine import unpack_terms
engine = aslsd_b200.Engine(0)
rng = np.random.default_rng(909)
T_f = 300.
sparse = np.sort(rng.uniform(0, T_f, size=700))
dense = rng.uniform(0, T_f, size=30000)
bursts = np.repeat(rng.uniform(5., T_f - 5., size=12), 40)
dense_b = np.sort(np.concatenate([dense, bursts, sparse[:30]]))
dense = np.sort(dense)
basis = [{'type': 'gauss'}]
def run(times, label, sigma=1.0, delta=0.0):
    d = len(times)
    model = {'d': d, 'kernels': [[basis for _ in range(d)] for _ in range(d)]}
    engine.set_events(times, T_f)
    engine.set_model(model)
    for k in range(d):
        x = np.array([0.6] + [0.2, sigma, delta, 0.1, 1.3 * sigma, 0.5 * delta][:3 * d])
        engine.use_lag_quadrature(True)
        a = engine.exact_sums_packed(k, x)
        wa = engine.last_work_counts()
        engine.use_lag_quadrature(False)
        b = engine.exact_sums_packed(k, x)
        engine.use_lag_quadrature(True)
        err = np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
        print(label, 'k', k, 'pairs', wa['moment_pairs'], 'streams', wa['streams'], 'bins', wa['bins'], 'max rel', err, flush=True)
run([sparse], 'sparse only')
run([dense[:5000]], 'dense 5000 only')
run([dense], 'dense only')
run([dense_b], 'dense + bursts only')
run([sparse, dense], 'sparse, dense')
run([sparse, dense_b], 'sparse, dense + bursts')
run([sparse, dense_b], 'sparse, dense + bursts, two passes', 0.4, 4.0)
