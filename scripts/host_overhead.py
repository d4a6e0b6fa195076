"""Host-side cost of one batched exact-sum call (planning, uploads, assembly) on a tiny stream."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import aslsd_b200
from aslsd_b200 import simulate
model, times, T_f, xs = simulate.config_c5(T_f=300.)
eng = aslsd_b200.Engine(0)
eng.set_events(times, T_f); eng.set_model(model)
d = model['d']
for rep in range(4):
    t0 = time.perf_counter(); eng.exact_sums_batch(list(range(d)), xs); t1 = time.perf_counter()
    print('batch call wall %.2f ms, kernels %.3f ms' % (1e3 * (t1 - t0), eng.last_timing()[0]))
import torch
raw = torch.zeros(eng.exact_raw_len_batch(d), dtype=torch.float64, device='cuda')
for rep in range(3):
    t0 = time.perf_counter(); eng.exact_raw_device_batch(list(range(d)), xs, raw.data_ptr()); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    h = raw.cpu().numpy(); t3 = time.perf_counter()
    eng.exact_assemble_batch(list(range(d)), xs, h); t4 = time.perf_counter()
    print('raw_device %.2f ms (+sync %.2f), d2h %.2f ms, assemble %.2f ms' % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)))
