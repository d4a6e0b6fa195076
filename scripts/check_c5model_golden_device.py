"""Device against the LIVE REFERENCE's sums of the benchmark's own model (tests/golden/exact_c5model.npz:
ten event types, SBF-Gaussian(10), sub-problems 0, 4 and 9 on a short path), gate 1e-10 per term."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import aslsd_b200  # noqa: E402
from aslsd_b200 import simulate  # noqa: E402
from helpers import RTOL, assert_sums_close, golden_times, load_golden  # noqa: E402

g = load_golden('exact_c5model')
times, T_f = golden_times(g)
engine = aslsd_b200.Engine(0)
engine.set_events(times, T_f)
engine.set_model(simulate.sbf_gauss_model(10, 10))
for k in (int(v) for v in g['ks']):
    sum_f, sum_diff_f = engine.exact_sums(k, g['k%d/x' % k])
    worst = assert_sums_close(sum_f, sum_diff_f, g, 'k%d' % k, rtol=RTOL)
    print('k = %d: %d terms, worst relative difference %.2e, work %s' % (k, len(sum_f), worst, engine.last_work_counts()),
          flush=True)
print('device equals the reference on the benchmark model')
