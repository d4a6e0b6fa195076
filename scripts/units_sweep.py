"""Tuning knobs of the moment kernel against its time on the C5 step:

    python scripts/units_sweep.py [ENV_NAME] value ...     # default ASLSD_LQ_UNITS (work-unit count);
                                                            # ASLSD_LQ_FILL: mean pairs of a tile / chunk capacity
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, json
sys.path.insert(0, %r)
import numpy as np, torch, aslsd_b200
from aslsd_b200 import simulate
model, times, T_f, xs = simulate.config_c5(T_f=1.1e6)
eng = aslsd_b200.Engine(0); eng.set_events(times, T_f); eng.set_model(model); eng.use_weight_cache(False)
ks = list(range(10)); flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda:0'); ms = []
for rep in range(7):
    flush.zero_(); torch.cuda.synchronize(); eng.exact_sums_batch(ks, xs)
    if rep >= 2: ms.append(eng.last_moment_timing())
im = min(eng.index_maps_kernel_time()[0] for _ in range(4))
print(json.dumps({'units': %r, 'moment_ms': float(np.mean(ms)), 'min': float(np.min(ms)), 'index_maps_ms': im}))
'''
args = sys.argv[1:]
name = 'ASLSD_LQ_UNITS'
if args and args[0].startswith('ASLSD_'):
    name, args = args[0], args[1:]
for units in args:
    env = dict(os.environ, **{name: units})
    r = subprocess.run([sys.executable, '-c', code % (ROOT, units)], env=env, capture_output=True, text=True)
    print(r.stdout.strip().split('\n')[-1] if r.returncode == 0 else r.stderr[-500:], flush=True)
