"""Phases of one sharded end-to-end C5 step under torchrun (rank 0 prints; max over ranks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import aslsd_b200
from aslsd_b200 import simulate
from aslsd_b200.distributed import ShardedExactSums, time_block_ranges, block_slices

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
model, times, T_f, xs = simulate.config_c5(d=10, n_basis=10, T_f=1.1e6, seed=1234)
times = [torch.from_numpy(t).pin_memory().numpy() for t in times]
d = 10
ks = list(range(d))
eng = aslsd_b200.Engine(local)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.set_events(times, T_f)
eng.set_model(model)
eng.use_weight_cache(False)
eng.use_async_events(True)
sh = ShardedExactSums(eng, times, rank=rank, world_size=world, dist=dist, device=dev)
acc = {}


def lap(label, t0):
    t1 = time.perf_counter()
    acc.setdefault(label, []).append(1e3 * (t1 - t0))
    return t1


for it in range(6):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = t = time.perf_counter()
    block = time_block_ranges(times, world)[rank]
    t = lap('time_block_ranges', t)
    sh.max_lag = eng.exact_max_lag(ks, xs)
    t = lap('exact_max_lag', t)
    local_t, ranges = block_slices(times, block, sh.max_lag)
    t = lap('block_slices', t)
    eng.set_events(local_t, T_f)
    sh.ranges = ranges
    t = lap('set_events (returns)', t)
    raw = sh.launch_batch(ks, xs)
    t = lap('launch_batch (plan + enqueue + allreduce enqueue)', t)
    host = raw.cpu().numpy()
    t = lap('raw.cpu() (waits for kernels + allreduce)', t)
    out = eng.exact_assemble_batch(ks, xs, host)
    t = lap('assemble_batch', t)
    lap('total', t0)
if True:
    for label, v in acc.items():
        x = torch.tensor([float(np.mean(v[1:]))], dtype=torch.float64, device=dev)
        mx = x.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        if rank == 0:
            print('%-52s rank0 %7.2f ms   max %7.2f ms' % (label, x.item(), mx.item()))
    if rank == 0:
        print('kernels %.2f ms' % eng.last_timing()[0])
dist.destroy_process_group()
