"""Kernel time and HBM fraction of the batched searchsorted at C5 size (10 types, ~10^7 events)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aslsd_b200
from aslsd_b200 import simulate

T_f = float(sys.argv[1]) if len(sys.argv) > 1 else 1.1e6
model, times, T_f, xs = simulate.config_c5(T_f=T_f)
eng = aslsd_b200.Engine(0)
eng.set_events(times, T_f)
best = None
for rep in range(5):
    ms, nbytes = eng.index_maps_kernel_time()
    best = ms if best is None else min(best, ms)
peak = 6540.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
print(json.dumps({'kernel_ms': best, 'algorithmic_bytes': nbytes, 'GBps': nbytes / best * 1e-6,
                  'frac_of_hbm_peak': nbytes / best * 1e-6 / peak}))
