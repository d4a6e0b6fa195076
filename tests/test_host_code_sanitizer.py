"""The HOST side of libaslsd_b200.so (argument checks, planning, packing, assembly, buffer
management behind every C-ABI entry point) under AddressSanitizer and
UndefinedBehaviorSanitizer, on a machine without a GPU.

The library's sources are built a second time against a do-nothing CUDA runtime
(tests/host_stub/cudart_stub.cpp: allocations are host memory, kernel launches are no-ops) and a
driver makes every public engine call on ten models (one to three event types, every basis
family, marked models) in a subprocess with the sanitizer runtime preloaded.  Nothing is computed -
this is not a CPU path and the product never loads that build; what fails the test is a memory
error or a wrong error code in the host code."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'host_stub'))


@pytest.mark.timeout(900)
def test_host_code_runs_clean_under_sanitizers():
    import build_stub
    runtimes = build_stub.sanitizer_runtimes()
    if runtimes is None:
        pytest.skip('no sanitizer runtimes on this machine')
    lib = build_stub.build()
    env = dict(os.environ, LD_PRELOAD=' '.join(runtimes), ASAN_OPTIONS='detect_leaks=0:abort_on_error=0',
               UBSAN_OPTIONS='print_stacktrace=1')
    res = subprocess.run([sys.executable, os.path.join(HERE, 'host_stub', 'drive_host_code.py'), lib],
                         capture_output=True, text=True, env=env, timeout=800)
    assert res.returncode == 0 and 'host code ok' in res.stdout, \
        (res.stdout[-2000:] + '\n' + res.stderr[-6000:])
    assert 'AddressSanitizer' not in res.stderr and 'runtime error:' not in res.stderr
