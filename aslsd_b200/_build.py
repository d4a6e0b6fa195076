"""In-tree build of libaslsd_b200.so with nvcc for sm_100a."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, '_lib')
LIB = os.path.join(LIB_DIR, 'libaslsd_b200.so')
SOURCES = ['ctx.cu', 'index_maps.cu', 'plan.cu', 'exact.cu', 'sampled.cu', 'sampler.cu', 'point_sums.cu']
HEADERS = ['common.cuh', 'basis_math.cuh', 'pair_kernels.cuh', 'pair_tiles.h', 'plan.h', 'lag_quadrature.cuh',
           os.path.join('..', '..', 'include', 'aslsd_b200.h')]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-O3', '-std=c++17', '-lineinfo',
           '-gencode', 'arch=compute_100a,code=sm_100a',
           '-Xcompiler', '-fPIC', '-Xptxas', '-v', '-shared', '-o', LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    with open(os.path.join(LIB_DIR, 'ptxas.log'), 'w') as f:
        f.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stderr[-4000:])
    if verbose:
        print(res.stderr)
    return LIB
