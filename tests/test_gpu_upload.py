"""Upload paths of aslsd_set_events: pageable arrays (staged and checked by host threads) and
page-locked arrays (uploaded in place, checked on the device, verdict now or deferred)."""
import numpy as np
import pytest

from cases import CASES, case_x

pytestmark = pytest.mark.gpu


def _pinned(arrays):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory().numpy() for a in arrays]


@pytest.fixture()
def engine():
    import aslsd_b200
    eng = aslsd_b200.Engine(0)
    yield eng
    eng.close()


def test_page_locked_inputs_give_identical_sums(engine):
    name = 'gauss2d'
    model = CASES[name][0]
    rng = np.random.default_rng(3)
    T_f = 5000.
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (7001, 5500)]   # odd length: 3 pad values
    engine.set_model(model)
    engine.set_events(times, T_f)
    want = [engine.exact_sums_packed(k, case_x(name)[k]) for k in range(2)]
    kap, var, lag = engine.index_maps()
    for async_events in (False, True):
        engine.use_async_events(async_events)
        pinned = _pinned(times)
        engine.set_events(pinned, T_f)
        got = [engine.exact_sums_packed(k, case_x(name)[k]) for k in range(2)]
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
        kap2, var2, lag2 = engine.index_maps()
        for i in range(2):
            for j in range(2):
                np.testing.assert_array_equal(kap2[i][j], kap[i][j])
                np.testing.assert_array_equal(var2[i][j], var[i][j])
        # a mix of page-locked and pageable arrays
        engine.set_events([pinned[0], times[1]], T_f)
        np.testing.assert_array_equal(engine.exact_sums_packed(1, case_x(name)[1]), want[1])


@pytest.mark.parametrize('async_events', [False, True])
@pytest.mark.parametrize('defect', ['swap', 'nan', 'first_nan'])
def test_unsorted_page_locked_inputs_are_rejected(engine, async_events, defect):
    import aslsd_b200
    name = 'exp2d'
    model = CASES[name][0]
    rng = np.random.default_rng(4)
    T_f = 3000.
    good = [np.sort(rng.uniform(0, T_f, size=n)) for n in (6000, 4000)]
    bad = [t.copy() for t in good]
    if defect == 'swap':
        bad[1][[2500, 2501]] = bad[1][[2501, 2500]]
    elif defect == 'nan':
        bad[0][1234] = np.nan
    else:
        bad[0][0] = np.nan
    engine.set_model(model)
    engine.use_async_events(async_events)
    # pageable: always rejected by set_events itself
    with pytest.raises(aslsd_b200.EngineError, match='sorted'):
        engine.set_events(bad, T_f)
    pinned = _pinned(bad)
    if not async_events:
        with pytest.raises(aslsd_b200.EngineError, match='sorted'):
            engine.set_events(pinned, T_f)
    else:
        engine.set_events(pinned, T_f)          # returns with the transfer in flight
        with pytest.raises(aslsd_b200.EngineError, match='sorted'):
            engine.exact_sums_packed(0, case_x(name)[0])
    # the context has no events afterwards, and takes good ones again
    with pytest.raises(aslsd_b200.EngineError):
        engine.exact_sums_packed(0, case_x(name)[0])
    engine.set_events(_pinned(good), T_f)
    a = engine.exact_sums_packed(0, case_x(name)[0])
    engine.set_events(good, T_f)
    np.testing.assert_array_equal(a, engine.exact_sums_packed(0, case_x(name)[0]))


def test_views_into_page_locked_arrays_are_uploaded_in_place(engine):
    """The time-block shards of ShardedExactSums.set_events are slices of the caller's arrays:
    interior pointers of a page-locked allocation, at odd offsets too."""
    import types
    from aslsd_b200.distributed import ShardedExactSums
    name = 'sbf2d3'
    model = CASES[name][0]
    rng = np.random.default_rng(9)
    T_f = 4000.
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (9001, 7000)]
    xs = case_x(name)
    engine.set_model(model)
    engine.set_events(times, T_f)
    want = engine.exact_sums_batch([0, 1], xs)
    engine.use_async_events(True)
    pinned = _pinned(times)
    # rank 1 of 3: its block plus the max-lag halo, cut out of the pinned arrays
    import torch
    # no collective: the shards are summed below
    fake_dist = types.SimpleNamespace(all_reduce=lambda *a, **k: None, ReduceOp=types.SimpleNamespace(SUM=None))
    parts = []
    for rank in range(3):
        sh = ShardedExactSums(engine, pinned, rank=rank, world_size=3, dist=fake_dist,
                              device=torch.device('cuda', 0))
        sh.set_events(pinned, T_f, [0, 1], xs)
        assert all(len(t) < len(p) for t, p in zip(engine._events_arrays, pinned))
        raw = sh.launch_batch([0, 1], xs)
        parts.append(raw.cpu().numpy().copy())
    total = parts[0] + parts[1] + parts[2]
    got = engine.exact_assemble_batch([0, 1], xs, total)
    for a, b in zip(got, want):
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)


def test_staged_moment_launches_give_identical_sums(engine):
    """Page-locked uploads of at least 24 MB go to the copy stream and the moment kernels of the next
    exact-sum call start in pieces, type by type as the data arrives, on two lanes (exact.cu:
    launch_moments).  The pieces cut the same work units out of the same launch: bit-identical sums,
    first call and repeated calls, against the pageable upload of the same events; other entry
    points used right after such an upload (index maps) are ordered behind it."""
    from aslsd_b200 import simulate
    d = 4
    model = simulate.sbf_gauss_model(d, 3)
    rng = np.random.default_rng(21)
    T_f = 9.0e5
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (900_001, 850_000, 910_000, 880_003)]   # 28 MB
    xs = [np.concatenate([[0.5], rng.uniform(0.02, 0.1, size=3 * d)]) for _ in range(d)]
    ks = list(range(d))
    engine.set_model(model)
    engine.use_weight_cache(False)
    engine.set_events(times, T_f)
    want = engine.exact_sums_batch(ks, xs)
    kap, var, lag = engine.index_maps(want_lag_sizes=False)
    pinned = _pinned(times)
    for async_events in (True, False):
        engine.use_async_events(async_events)
        for rep in range(3):
            engine.set_events(pinned, T_f)
            got = engine.exact_sums_batch(ks, xs)
            for a, b in zip(got, want):
                np.testing.assert_array_equal(a, b)
        again = engine.exact_sums_batch(ks, xs)          # events resident: one launch
        for a, b in zip(again, want):
            np.testing.assert_array_equal(a, b)
        engine.set_events(pinned, T_f)
        kap2, var2, _ = engine.index_maps(want_lag_sizes=False)   # straight after the upload
        for i in range(d):
            for j in range(d):
                np.testing.assert_array_equal(kap2[i][j], kap[i][j])
                np.testing.assert_array_equal(var2[i][j], var[i][j])
