"""Planning and assembly - the host half of an exact-sum call - against the oracle, without a GPU.

Which parameters and event streams enter which row of the raw accumulator array
(aslsd_b200/csrc/plan.cu, exact.cu) and how the rows combine into the packed per-term sums
(csrc/plan.h) is everything in the exact path that depends on the number of event types; the
device kernels only traverse the event pairs of a row.  The driver
(tests/host_stub/check_plan_assembly.py) plans a batch with the library built against the
do-nothing CUDA runtime, reads the rows back (aslsd_exact_plan_rows), fills them with numpy from the
component definitions, and lets the library assemble: values and derivatives must equal the
literal oracle's for models with one to five event types and for the benchmark's own ten-type model
(exponential kernels, free Gaussians,
sums of Gaussians with fixed shapes - the family of configuration C5 -, mixed fixed / free patterns
and kernels with an exponential and a Gaussian basis, every derivative included).  The same driver
checks the truncation a plan prescribes (lag windows, limits at infinity near T_f) over random
parameters: rows evaluated in full and the way the kernels traverse them give the same packed sums;
and the plans of time-block shards (two to five event types): every row is restricted to the block's
targets of its own target type, the packed sums of the blocks add up to those of the whole path.
Marked models run end to end: the B200 estimator over the library's planning, role / mask handling and
assembly, the weighted rows evaluated with numpy and handed back through a substituted read-back,
against the golden vectors of the reference's MTLHExactEstim (eight cases, one to three event types)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'host_stub'))


@pytest.mark.timeout(900)
def test_rows_and_assembly_match_the_oracle():
    import build_stub
    runtimes = build_stub.sanitizer_runtimes()
    if runtimes is None:
        pytest.skip('no sanitizer runtimes on this machine')
    lib = build_stub.build()
    env = dict(os.environ, LD_PRELOAD=' '.join(runtimes), ASAN_OPTIONS='detect_leaks=0:abort_on_error=0',
               UBSAN_OPTIONS='print_stacktrace=1')
    res = subprocess.run([sys.executable, os.path.join(HERE, 'host_stub', 'check_plan_assembly.py'), lib],
                         capture_output=True, text=True, env=env, timeout=800)
    assert res.returncode == 0 and 'plan and assembly ok' in res.stdout, \
        (res.stdout[-3000:] + '\n' + res.stderr[-6000:])
    assert 'AddressSanitizer' not in res.stderr and 'runtime error:' not in res.stderr
    assert 'truncation: 24 parameter draws' in res.stdout
    assert res.stdout.count('marked model: estimator state equals the golden vectors') == 8   # incl. two with
    # baselines constant on intervals (psi sums at moved horizons)
    assert res.stdout.count('three time blocks add up to the whole path') == 4
    assert res.stdout.count('derivatives=True') == 19 and res.stdout.count(' d=5 ') == 2
    assert 'C5       d=10 rows=11440' in res.stdout      # the benchmark's own model: the row count of bench.py
