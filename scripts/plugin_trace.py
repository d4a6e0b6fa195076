"""Per-call host phases of the plug-in path (d x MHPExactEstimB200.lse_k_grad_estimate at C5):
run with ASLSD_TRACE=1 to see the library's own phase timers on stderr."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import aslsd_b200
from aslsd_b200 import simulate

DIRECT = os.environ.get('PLUGIN_TRACE_DIRECT', '1') == '1'   # a direct engine call before every estimator call
model, times, T_f, xs = simulate.config_c5(T_f=float(sys.argv[1]) if len(sys.argv) > 1 else 1.1e6)
engine = aslsd_b200.Engine(0)
path = aslsd_b200.ProcessPath(times, T_f, book_keeping=False, engine=engine)
mirror = aslsd_b200.MHP(model)
d = model['d']
ests = []
for k in range(d):
    est = aslsd_b200.MHPExactEstimB200(engine=engine, reference_recombine=False)
    est.initialize(k, 4, mirror, path)
    ests.append(est)
for cache in (False, True):
    engine.use_weight_cache(cache)
    for rep in range(3):
        engine._events_key = None
        engine.bind(path, mirror)
        for k in range(d):
            t0 = time.perf_counter()
            x = np.asarray(xs[k], dtype=np.float64)
            if DIRECT:
                packed = engine.exact_sums_packed(k, x, False, True)
            t1 = time.perf_counter()
            g = ests[k].lse_k_grad_estimate(xs[k])
            t2 = time.perf_counter()
            print('cache=%d rep=%d k=%d  engine call %.2f ms (kernels %.2f, moment %.2f)  estimator call %.2f ms'
                  % (cache, rep, k, 1e3 * (t1 - t0), engine.last_timing()[0], engine.last_moment_timing(),
                     1e3 * (t2 - t1)), flush=True)
