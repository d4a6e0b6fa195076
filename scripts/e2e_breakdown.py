"""Host-side breakdown of one end-to-end C5 step (GPU box): where the non-kernel time goes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import aslsd_b200
from aslsd_b200 import simulate
from aslsd_b200.distributed import ShardedExactSums

T_f = float(sys.argv[1]) if len(sys.argv) > 1 else 1.1e6
model, times, T_f, xs = simulate.config_c5(d=10, n_basis=10, T_f=T_f, seed=1234)
d = 10
eng = aslsd_b200.Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.set_events(times, T_f)
eng.set_model(model)
ks = list(range(d))
sh = ShardedExactSums(eng, times, device=torch.device('cuda', 0))
sh.packed_batch(ks, xs)
torch.cuda.synchronize()


def tick(label, fn, n=3):
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print('%-46s %8.2f ms' % (label, 1e3 * best))
    return out


tick('set_events', lambda: eng.set_events(times, T_f))
raw = tick('raw_device_batch after set_events (re-plan)', lambda: (eng.set_events(times, T_f), sh.launch_batch(ks, xs))[1])
raw = tick('raw_device_batch, plan cached', lambda: sh.launch_batch(ks, xs))
print('   kernels %.2f ms, %d launches' % eng.last_timing())
host = tick('raw.cpu().numpy()', lambda: raw.cpu().numpy())
tick('assemble_batch', lambda: eng.exact_assemble_batch(ks, xs, host))
tick('packed_batch (all, plan cached)', lambda: sh.packed_batch(ks, xs))
tick('exact_sums_batch (C ABI, host buffers)', lambda: eng.exact_sums_batch(ks, xs))
tick('set_events + exact_sums_batch', lambda: (eng.set_events(times, T_f), eng.exact_sums_batch(ks, xs)))
