"""In-tree build of libaslsd_b200.so with nvcc for sm_100a (one object per source, in parallel)."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, '_lib')
LIB = os.path.join(LIB_DIR, 'libaslsd_b200.so')
SOURCES = ['ctx.cu', 'index_maps.cu', 'plan.cu', 'exact.cu', 'sampled.cu', 'sampler.cu', 'point_sums.cu',
           'simulate.cu']
HEADERS = ['common.cuh', 'basis_math.cuh', 'pair_kernels.cuh', 'pair_tiles.h', 'plan.h', 'lag_quadrature.cuh',
           os.path.join('..', '..', 'include', 'aslsd_b200.h')]
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _mtime(path):
    return os.path.getmtime(path) if os.path.exists(path) else 0.


def needs_build(lib=LIB):
    t = _mtime(lib)
    if not t:
        return True
    return any(_mtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, defines=(), lib=LIB):
    """Compile every source that exists in csrc/ (``defines``: extra -D flags, used by the kernel
    variant experiments of scripts/) and link ``lib``."""
    if not force and not needs_build(lib):
        return lib
    os.makedirs(LIB_DIR, exist_ok=True)
    tag = os.path.splitext(os.path.basename(lib))[0]
    obj_dir = os.path.join(LIB_DIR, 'obj_' + tag)
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    sources = [f for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    newest_header = max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)
    stamp = os.path.join(obj_dir, 'defines.txt')
    same_defines = os.path.exists(stamp) and open(stamp).read() == ' '.join(defines)

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        if (not force and same_defines and _mtime(obj) > max(_mtime(os.path.join(CSRC, src)), newest_header)):
            return obj, ''
        cmd = [nvcc] + FLAGS + list(defines) + ['-c', src, '-o', obj]
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, res.stderr[-4000:]))
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as pool:
        results = list(pool.map(compile_one, sources))
    with open(stamp, 'w') as f:
        f.write(' '.join(defines))
    log = ''.join(r[1] for r in results)
    if log:
        with open(os.path.join(LIB_DIR, 'ptxas_%s.log' % tag), 'w') as f:
            f.write(log)
    res = subprocess.run([nvcc, '-shared', '-o', lib] + [r[0] for r in results], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('link failed:\n' + res.stderr[-4000:])
    if verbose:
        print(log)
    return lib
