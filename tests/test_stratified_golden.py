"""Host-side ASLSD logic against the reference's golden traces (CPU).

The estimator, stratification, sampling and solver mirrors of aslsd_b200 are
driven with the oracle test double in place of the device
(tests/oracle_engine.py): what is pinned here is the host logic and the
stratified oracle itself; tests/test_gpu_stratified.py replays the same traces
through the CUDA engine."""
import numpy as np
import pytest

from helpers import replay_strat_golden
from oracle_engine import OracleEngine

NAMES = ['strat_exp1d', 'strat_exp2d', 'strat_exp2d_gradalloc', 'strat_gauss1d', 'strat_sbf2d3', 'strat_mix1d']


@pytest.mark.parametrize('name', NAMES)
def test_replay_reference_trace(name):
    replay_strat_golden(name, OracleEngine(), mode='strict', free_running=(name == 'strat_exp1d'))
    replay_strat_golden(name, OracleEngine(), mode='forced')


@pytest.mark.parametrize('name', ['strat_gauss1d_long', 'strat_sbf2d3_long', 'strat_exp2d_long'])
def test_replay_long_reference_trace(name):
    """100 - 120 iterations under grad_alloc=False (the fit default): every integer allocation,
    hence the whole RNG stream, is the reference's."""
    replay_strat_golden(name, OracleEngine(), mode='strict')


def test_replay_reference_trace_of_the_benchmark_model():
    """The benchmark's own model (config C5: ten event types, SBF-Gaussian(10)) under the reference's
    stratified estimator (tests/golden/make_golden_strat.py, 'strat_c5model': three ADAM iterations of
    every sub-problem on a 2 255-event path).  The oracle double takes ~25 s per estimator call at
    d = 10, so the default run replays the first call of sub-problem 0 - 121 terms, whose sums can
    only match if every sample of every term was drawn in the reference's order - and the whole trace
    (30 calls, final generator state included) with ASLSD_FULL_REPLAY=1: profiles/r2_strat_c5model_replay.log."""
    import os
    full = os.environ.get('ASLSD_FULL_REPLAY') == '1'
    worst = replay_strat_golden('strat_c5model', OracleEngine(), mode='strict', first_calls=None if full else 1)
    print('strat_c5model %s replay: worst relative difference of a per-term sum %.2e' % ('full' if full else 'first-call', worst))


def test_stratification_defaults_match_survey_geometry():
    """Default geometry at N = 10^6 (measured on the reference)."""
    import aslsd_b200
    n = 1_000_000
    s = aslsd_b200.GeneralStratification(n=n)
    assert s.n_exact == 1000 and s.exact_domain[0] == n - 1000 and s.exact_domain[-1] == n - 1
    assert s.n_adaptive_strata == 10 and s.adaptive_strata[-1][1] == n - 1001
    assert list(s.adaptive_strata_sizes) == [1000] * 9 + [1001]
    assert s.n_nonadaptive_strata == 10 and s.nonadaptive_strata[0][0] == 0
    assert s.nonadaptive_strata[-1][1] == s.adaptive_strata[0][0] - 1
    assert list(s.nonadaptive_abs_alloc) == [1000] * 10
    assert s.ema_weight == .6 and np.isclose(s.adaptive_rel_alloc.sum(), 1.)
    dstr = aslsd_b200.DoubleIxStratification(h_max=999_999)
    assert [list(h) for h in dstr.adaptive_strata] == [[h, h] for h in range(1, 30)]
    assert dstr.nonadaptive_strata[0][0] == 30 and dstr.nonadaptive_strata[-1][1] == 999_999
    assert len(dstr.nonadaptive_strata) == 9
    lag_sizes = [[np.concatenate([[0], n - np.arange(1, n)])]]
    dstr.get_stratification_size(lag_sizes, 0, 0)
    assert list(dstr.nonadaptive_abs_alloc) == [1111] * 9
    assert dstr.n_samples_adaptive[0] == 10 ** 4


def test_estimator_pickle_drops_device_handle():
    import pickle
    import aslsd_b200
    est = aslsd_b200.MHPStratEstimB200(engine=OracleEngine())
    clone = pickle.loads(pickle.dumps(est))
    assert clone.engine is None and clone.t == 0
    with pytest.raises(RuntimeError):
        clone._bound_engine()


def test_numpy_summation_order_is_the_one_restated():
    """np.sum / np.mean / np.std of a contiguous float64 array follow the pairwise
    recursion that oracle.stratified.numpy_pairwise_sum (and the CUDA stat
    kernel) restate: bit-for-bit on random data of many lengths."""
    from oracle import stratified as OS
    rng = np.random.default_rng(3)
    for n in list(range(1, 40)) + [127, 128, 129, 255, 256, 257, 345, 1000, 1111, 4097, 10001]:
        a = rng.normal(size=n) * 10. ** rng.integers(-3, 4, size=n)
        assert OS.numpy_pairwise_sum(a) == np.sum(a), n
        mean = OS.numpy_pairwise_sum(a) / n
        assert mean == np.mean(a)
        dev = (a - mean) * (a - mean)
        assert np.sqrt(OS.numpy_pairwise_sum(dev) / n) == np.std(a), n
