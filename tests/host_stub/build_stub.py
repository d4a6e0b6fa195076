"""Build (tests/_hoststub/, git-ignored) the do-nothing CUDA runtime and a copy of the library's
sources linked against it with AddressSanitizer.  Test infrastructure only - see cudart_stub.cpp."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(os.path.dirname(HERE), '_hoststub')
CSRC = os.path.join(ROOT, 'aslsd_b200', 'csrc')
STUB = os.path.join(OUT, 'libcudart_stub.so')
LIB = os.path.join(OUT, 'libaslsd_b200_hoststub.so')


def sanitizer_runtimes():
    """Paths of the AddressSanitizer and UndefinedBehaviorSanitizer runtimes to preload, or None."""
    paths = []
    for name in ('libasan.so', 'libubsan.so'):
        path = subprocess.run(['gcc', '-print-file-name=' + name], capture_output=True, text=True).stdout.strip()
        if not (os.path.isabs(path) and os.path.exists(path)):
            return None
        paths.append(path)
    return paths


def build():
    from aslsd_b200 import _build
    os.makedirs(OUT, exist_ok=True)
    sources = [os.path.join(CSRC, f) for f in _build.SOURCES]
    deps = sources + [os.path.join(CSRC, h) for h in _build.HEADERS] + [os.path.join(HERE, 'cudart_stub.cpp')]
    if os.path.exists(LIB) and os.path.exists(STUB) and \
            all(os.path.getmtime(f) <= os.path.getmtime(LIB) for f in deps):
        return LIB
    cuda = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    subprocess.run(['g++', '-O1', '-g', '-fPIC', '-shared', '-fsanitize=address,undefined', '-I', os.path.join(cuda, 'include'),
                    '-o', STUB, os.path.join(HERE, 'cudart_stub.cpp')], check=True, capture_output=True)
    nvcc = os.environ.get('NVCC', os.path.join(cuda, 'bin', 'nvcc'))
    res = subprocess.run([nvcc, '-O1', '-g', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
                          '-Xcompiler', '-fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer', '--cudart', 'none',
                          '-shared', '-o', LIB] + _build.SOURCES +
                         ['-L' + OUT, '-lcudart_stub', '-lubsan', '-Xlinker', '-rpath,' + OUT],
                         cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('stub build failed:\n' + res.stderr[-3000:])
    return LIB
