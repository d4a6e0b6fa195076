"""Host phases of one sharded end-to-end C5 step on the packed path (what bench.py's e2e times at N > 1),
measured on ONE GPU that plays rank R of W (no communication: the all-reduce of 11 110 doubles is
left out, everything else is what the rank does).

    python scripts/e2e_packed_breakdown.py [W] [R]        # default 8 3
    ASLSD_TRACE=1 python scripts/e2e_packed_breakdown.py  # + the library's own phase timers on stderr
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import aslsd_b200  # noqa: E402
from aslsd_b200 import simulate  # noqa: E402
from aslsd_b200.distributed import ShardedExactSums, block_slices  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
R = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device('cuda', 0)
model, times, T_f, xs = simulate.config_c5(d=10, n_basis=10, T_f=1.1e6, seed=1234)
times = [torch.from_numpy(t).pin_memory().numpy() for t in times]
d = 10
ks = list(range(d))
eng = aslsd_b200.Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.set_events(times, T_f)
eng.set_model(model)
eng.use_weight_cache(False)
eng.use_async_events(True)
sh = ShardedExactSums(eng, times, rank=R, world_size=W, dist=None, device=dev)
sh.world_size = 1          # no all-reduce; the blocks below are those of rank R of W
acc = {}


def lap(label, t0):
    t1 = time.perf_counter()
    acc.setdefault(label, []).append(1e3 * (t1 - t0))
    return t1


from aslsd_b200.distributed import time_block_ranges  # noqa: E402
blocks = time_block_ranges(times, W)
for it in range(8):
    torch.cuda.synchronize()
    t0 = t = time.perf_counter()
    block = blocks[R]                                   # cached by ShardedExactSums._blocks for the same arrays
    sh.max_lag = 1.25 * eng.exact_max_lag(ks, xs)
    t = lap('exact_max_lag', t)
    local_t, ranges = block_slices(times, block, sh.max_lag)
    t = lap('block_slices', t)
    eng.set_events(local_t, T_f)
    sh.ranges = ranges
    t = lap('set_events (returns)', t)
    packed = sh.launch_packed_batch(ks, xs)
    t = lap('launch_packed_batch (plan + enqueue)', t)
    host = packed.cpu().numpy()
    t = lap('packed.cpu() (waits for the kernels)', t)
    out = eng.split_packed(ks, host)
    t = lap('split_packed', t)
    lap('total', t0)
for label, v in acc.items():
    print('%-44s %7.3f ms' % (label, float(np.mean(v[2:]))))
print('kernels %.3f ms, moment kernel %.3f ms' % (eng.last_timing()[0], eng.last_moment_timing()))
