"""Host pass vs host->device copy of one C5 upload, pageable and page-locked inputs (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import aslsd_b200
from aslsd_b200 import simulate

model, times, T_f, xs = simulate.config_c5(d=10, n_basis=10, T_f=1.1e6, seed=1234)
pinned = [torch.from_numpy(t).pin_memory().numpy() for t in times]
eng = aslsd_b200.Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.set_model(model)
ks = list(range(10))
for label, arrs in (('pageable', times), ('page-locked', pinned), ('pl + async', pinned)):
    eng.use_async_events(label == 'pl + async')
    for _ in range(3):
        eng.set_events(arrs, T_f)
        torch.cuda.synchronize()
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.set_events(arrs, T_f)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print('%-12s set_events returns after %.2f ms, copies done after %.2f ms' % (label, 1e3 * (t1 - t0), 1e3 * (t2 - t0)))
    eng.exact_sums_batch(ks, xs)
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.set_events(arrs, T_f)
        out = eng.exact_sums_batch(ks, xs)
        torch.cuda.synchronize()
        print('%-12s set_events + exact_sums_batch %.2f ms (kernels %.2f ms)' % (label, 1e3 * (time.perf_counter() - t0), eng.last_timing()[0]))
    ref = out if label == 'pageable' else ref
    if label != 'pageable':
        for a, b in zip(out[0] if isinstance(out, tuple) else out, ref[0] if isinstance(ref, tuple) else ref):
            assert np.array_equal(np.asarray(a), np.asarray(b))
        print('identical results')
