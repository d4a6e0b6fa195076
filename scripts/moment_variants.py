"""Build-time variants of the lag-moment kernel side by side (kernel experiments).

    python scripts/moment_variants.py build  name=-DFLAG,-DFLAG ...   # here (nvcc, no GPU needed)
    python scripts/moment_variants.py run [T_f]                       # on the GPU box: every built variant

``run`` loads each aslsd_b200/_lib/variants/lib_<name>.so in a subprocess, times the batched C5 step
(weight cache off, L2 flushed) and compares the packed sums with those of the first variant."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, 'aslsd_b200', '_lib', 'variants')


def build(specs):
    from aslsd_b200 import _build
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        name, _, flags = spec.partition('=')
        defines = [f for f in flags.split(',') if f]
        lib = os.path.join(VDIR, 'lib_%s.so' % name)
        _build.build(force=True, defines=defines, lib=lib)
        log = open(os.path.join(ROOT, 'aslsd_b200', '_lib', 'ptxas_lib_%s.log' % name)).read().split('\n')
        for i, line in enumerate(log):
            if 'lag_moment_kernelILb0' in line and 'Compiling' in line:
                print(name, ' | '.join(x.strip() for x in log[i + 1:i + 4]))


def run_one(lib, T_f):
    from aslsd_b200 import capi
    capi.LIB_PATH = lib
    import numpy as np
    import torch
    import aslsd_b200
    from aslsd_b200 import simulate
    model, times, T_f, xs = simulate.config_c5(T_f=T_f)
    eng = aslsd_b200.Engine(0)
    eng.set_events(times, T_f)
    eng.set_model(model)
    eng.use_weight_cache(False)
    d = model['d']
    ks = list(range(d))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda:0')
    ms, tot = [], []
    for rep in range(8):
        flush.zero_()
        torch.cuda.synchronize()
        out = eng.exact_sums_batch(ks, xs)
        if rep >= 3:
            ms.append(eng.last_moment_timing())
            tot.append(eng.last_timing()[0])
    np.save(lib + '.npy', np.concatenate(out))
    pairs = eng.last_work_counts()['moment_pairs']
    im = min(eng.index_maps_kernel_time()[0] for _ in range(4))
    print(json.dumps({'lib': os.path.basename(lib), 'moment_ms': float(np.mean(ms)), 'min': float(np.min(ms)),
                      'all_kernels_ms': float(np.mean(tot)), 'pairs': pairs, 'index_maps_ms': im}))


if __name__ == '__main__':
    if sys.argv[1] == 'build':
        build(sys.argv[2:])
    elif sys.argv[1] == 'one':
        run_one(sys.argv[2], float(sys.argv[3]))
    else:
        import numpy as np
        T_f = sys.argv[2] if len(sys.argv) > 2 else '1.1e6'
        libs = sorted(f for f in os.listdir(VDIR) if f.endswith('.so'))
        ref = None
        for f in libs:
            lib = os.path.join(VDIR, f)
            res = subprocess.run([sys.executable, os.path.abspath(__file__), 'one', lib, T_f], capture_output=True,
                                 text=True, timeout=600)
            line = res.stdout.strip().split('\n')[-1] if res.returncode == 0 else 'FAILED ' + res.stderr[-600:]
            if res.returncode == 0:
                got = np.load(lib + '.npy')
                if ref is None:
                    ref = got
                line += ' max_rel_vs_first=%.2e' % float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
            print(line, flush=True)
