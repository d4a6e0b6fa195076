"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# relative tolerance of the LSE path stated by BASELINE.json north_star
RTOL = 1e-10


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))


def golden_times(g):
    d = int(g['d'])
    return [np.asarray(g['times/%d' % i]) for i in range(d)], float(g['T_f'])


def iter_sum_entries(sum_f, sum_diff_f):
    """Yield (label, value) for every scalar / vector of a per-term result."""
    for name in sum_f:
        yield 'f/' + name, np.float64(sum_f[name])
        dv = sum_diff_f[name]
        if isinstance(dv, list):
            yield 'df1/' + name, np.asarray(dv[1], float)
            yield 'df2/' + name, np.asarray(dv[2], float)
        else:
            yield 'df/' + name, np.asarray(dv, float)


def assert_sums_close(sum_f, sum_diff_f, golden, prefix, rtol=RTOL, atol=0.0):
    """Per-term parity gate: |got - ref| <= rtol * |ref| (+ atol)."""
    worst = 0.0
    for label, got in iter_sum_entries(sum_f, sum_diff_f):
        ref = golden['%s/%s' % (prefix, label)]
        err = np.abs(np.asarray(got) - ref)
        bound = rtol * np.abs(ref) + atol
        assert np.all(err <= bound), (
            '%s %s: got %r ref %r (rel err %.3e)' % (
                prefix, label, got, ref,
                float(np.max(err / np.maximum(np.abs(ref), 1e-300)))))
        worst = max(worst, float(np.max(err / np.maximum(np.abs(ref), 1e-300))))
    return worst
