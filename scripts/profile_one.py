"""One batched exact-sum call on a reduced C5 stream (for ncu captures)."""
import sys
sys.path.insert(0, '.')
import aslsd_b200
from aslsd_b200 import simulate

T_f = float(sys.argv[1]) if len(sys.argv) > 1 else 5e4
model, times, T_f, xs = simulate.config_c5(T_f=T_f)
eng = aslsd_b200.Engine(0)
eng.set_events(times, T_f)
eng.set_model(model)
eng.use_weight_cache(False)     # every call runs the moment kernel
d = model['d']
for rep in range(3):
    eng.exact_sums_batch(list(range(d)), xs)
    print('rep', rep, eng.last_timing(), eng.last_work_counts())
