"""Timing probe of the exact sums on config C5 (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import aslsd_b200
from aslsd_b200 import simulate

T_f = float(sys.argv[1]) if len(sys.argv) > 1 else 1e5
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t0 = time.time()
model, times, T_f, xs = simulate.config_c5(T_f=T_f)
print('simulate %.1fs, events %d' % (time.time() - t0, sum(len(t) for t in times)))
eng = aslsd_b200.Engine(0)
t0 = time.time(); eng.set_events(times, T_f); print('set_events %.3fs' % (time.time() - t0))
eng.set_model(model)
t0 = time.time(); kap, var, lag = eng.index_maps(); print('index_maps %.3fs kernel %.3f ms' % (time.time() - t0, eng.last_timing()[0]))
del kap, var, lag
for k in range(nk):
    for rep in range(2):
        t0 = time.time()
        out = eng.exact_sums_packed(k, xs[k])
        dt = time.time() - t0
        ms, nl = eng.last_timing()
        print('k=%d rep=%d wall %.3fs kernels %.3f ms launches %d' % (k, rep, dt, ms, nl))
out, counts = eng.exact_sums_packed(0, xs[0], count_pairs=True)
tot = sum(counts.values())
print('pair terms k=0: %.4g  -> %.4g pair-terms/s' % (tot, tot / (ms * 1e-3)))
