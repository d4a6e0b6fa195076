"""The MTLH oracle (oracle/mtlh_exact.py) against the golden vectors generated from the
reference's MTLHExactEstim (tests/golden/make_golden_mtlh.py): every per-term sum, every
derivative block, M term, LSE and gradient."""
import numpy as np
import pytest

from cases import MTLH_CASES, mtlh_inputs
from helpers import load_golden
from oracle import mtlh_exact as OM


@pytest.mark.parametrize('name', sorted(MTLH_CASES))
def test_mtlh_oracle_matches_reference_golden(name):
    model, xs, n_events, T_f = MTLH_CASES[name]
    g = load_golden(name)
    times, marks, T_f = mtlh_inputs(name)
    for i in range(model['d']):
        np.testing.assert_array_equal(times[i], g['times/%d' % i])
        np.testing.assert_array_equal(marks[i], g['marks/%d' % i])
    for k in range(model['d']):
        x = g['k%d/x' % k]
        M, dM, sum_f, sum_diff_f = OM.exact_sums_literal(model, k, x, times, marks, T_f)
        flat = OM.flatten(M, dM, sum_f, sum_diff_f, 'k%d' % k)
        for key, val in flat.items():
            np.testing.assert_allclose(val, g[key], rtol=1e-13, atol=1e-13, err_msg=key)
        lse = OM.recombine_lse_k(model, k, M, sum_f, T_f)
        np.testing.assert_allclose(lse, g['k%d/lse' % k], rtol=1e-13)
        if 'k%d/grad' % k in g:
            grad = OM.recombine_grad_lse_k(model, k, dM, sum_diff_f, T_f)
            np.testing.assert_allclose(grad, g['k%d/grad' % k], rtol=1e-12, atol=1e-13)
