"""Replay the reference's stratified-estimator trace of the benchmark's own model (ten event types,
tests/golden/strat_c5model.npz) through MHPStratEstimB200 on the device, nothing forced but the
evaluation points: per-term sums, gradients, allocations and the final generator state must be the
reference's (the CPU suite replays the same trace over the oracle double)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

t0 = time.time()
import aslsd_b200  # noqa: E402
from helpers import replay_strat_golden  # noqa: E402

engine = aslsd_b200.Engine(0)
print('engine up %.1f s' % (time.time() - t0), flush=True)
t0 = time.time()
worst = replay_strat_golden('strat_c5model', engine, mode='strict', first_calls=2)
print('first two calls of sub-problem 0: worst relative difference %.2e (%.1f s)' % (worst, time.time() - t0), flush=True)
t0 = time.time()
worst = replay_strat_golden('strat_c5model', engine, mode='strict')
print('full trace (30 calls, final generator state): worst relative difference %.2e (%.1f s)' % (worst, time.time() - t0),
      flush=True)
