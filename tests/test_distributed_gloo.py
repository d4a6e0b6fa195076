"""Time-block sharding over 2 ranks on CPU (gloo): block boundaries, the
all-reduce of the per-shard accumulators and the assembly reproduce the
single-rank sums.  The device is replaced by the oracle test double; the same
code path runs with NCCL and the CUDA engine in bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import CASES, case_x


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, times, T_f, name, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here, os.path.join(here, 'golden')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import types
    from oracle_engine import OracleEngine
    from aslsd_b200.distributed import ShardedExactSums
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    engine = OracleEngine()
    engine.bind(types.SimpleNamespace(list_times=times, T_f=T_f), CASES[name][0])
    sharded = ShardedExactSums(engine, times, rank=rank, world_size=world, dist=dist,
                               device=torch.device('cpu'))
    x = case_x(name)[1]
    packed = sharded.packed(1, x)
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank), packed)
    both = sharded.packed_batch([0, 1], case_x(name))
    np.save(os.path.join(out_dir, 'batch%d.npy' % rank), np.concatenate(both))
    np.save(os.path.join(out_dir, 'ranges%d.npy' % rank), np.array(sharded.ranges))
    # second mode: every rank holds only its block plus the halo of the largest lag
    local_ranges = sharded.set_events(times, T_f, [0, 1], case_x(name))
    sliced = sharded.packed_batch([0, 1], case_x(name))
    np.save(os.path.join(out_dir, 'sliced%d.npy' % rank), np.concatenate(sliced))
    np.save(os.path.join(out_dir, 'held%d.npy' % rank), np.array([len(t) for t in engine.times]))
    # a parameter update that widens the kernels' reach beyond the uploaded halo (exponential
    # kernels with free decay rates): the launch re-slices instead of silently dropping sources
    engine2 = OracleEngine()
    engine2.bind(types.SimpleNamespace(list_times=times, T_f=T_f), CASES[WIDE_CASE][0])
    sharded2 = ShardedExactSums(engine2, times, rank=rank, world_size=world, dist=dist,
                                device=torch.device('cpu'))
    sharded2.set_events(times, T_f, [0, 1], case_x(WIDE_CASE), slack=1.0)
    lag_before = sharded2.max_lag
    again = sharded2.packed_batch([0, 1], _wide_x(WIDE_CASE))
    assert sharded2.max_lag > lag_before
    np.save(os.path.join(out_dir, 'wide%d.npy' % rank), np.concatenate(again))
    dist.barrier()
    dist.destroy_process_group()


WIDE_CASE = 'exp2d'


def _wide_x(name):
    """Evaluation point of ``name`` with every free kernel parameter but omega (first of each basis)
    scaled so that the lag windows grow by ~60 %."""
    from oracle import exact as OE
    from oracle import functionals as F
    model = CASES[name][0]
    out = []
    for k, x in enumerate(case_x(name)):
        x = np.array(x, dtype=float)
        for i, (a, b) in enumerate(OE.interval_map(model, k)):
            pos = a
            for basis in model['kernels'][k][i]:
                free = [v for v in range(2 if basis['type'] == 'exp' else 3) if v not in basis.get('fixed', {})]
                for v in free:
                    if v == 1:      # beta: decay rate (exp) or width (gauss)
                        x[pos] = x[pos] / 1.6 if basis['type'] == 'exp' else x[pos] * 1.6
                    pos += 1
        out.append(x)
    return out


def test_time_block_ranges_partition_the_events():
    from aslsd_b200.distributed import time_block_ranges
    rng = np.random.default_rng(0)
    times = [np.sort(rng.uniform(0, 100, size=n)) for n in (1000, 400, 2500)]
    for world in (1, 2, 3, 8):
        ranges = time_block_ranges(times, world)
        assert len(ranges) == world
        for i, t in enumerate(times):
            assert ranges[0][i][0] == 0 and ranges[-1][i][1] == len(t)
            for g in range(world - 1):
                assert ranges[g][i][1] == ranges[g + 1][i][0]
        # same cut time for every type, balanced target counts
        loads = [sum(b - a for a, b in r) for r in ranges]
        # every cut is within max(d, n / 2^16) events of its quantile
        assert max(loads) - min(loads) <= 2 * len(times)
        for g in range(world - 1):
            cut_lo = max(times[i][ranges[g][i][1] - 1] for i in range(3) if ranges[g][i][1] > 0)
            cut_hi = min(times[i][ranges[g][i][1]] for i in range(3) if ranges[g][i][1] < len(times[i]))
            assert cut_lo < cut_hi


@pytest.mark.timeout(300)
def test_two_rank_shards_allreduce_to_the_full_sums(tmp_path):
    from oracle import exact as OE
    name = 'sbf2d3'
    rng = np.random.default_rng(8)
    T_f = 600.
    times = [np.sort(rng.uniform(0, T_f, size=700)), np.sort(rng.uniform(0, T_f, size=500))]
    port = _free_port()
    mp.spawn(_worker, args=(2, port, times, T_f, name, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(os.path.join(tmp_path, 'rank%d.npy' % r)) for r in range(2)]
    np.testing.assert_array_equal(got[0], got[1])          # every rank holds the combined sums
    from oracle_engine import OracleEngine
    import types
    full = OracleEngine()
    full.bind(types.SimpleNamespace(list_times=times, T_f=T_f), CASES[name][0])
    want = full.exact_sums_packed(1, case_x(name)[1])
    np.testing.assert_allclose(got[0], want, rtol=1e-12, atol=1e-13)
    batch = np.load(os.path.join(tmp_path, 'batch0.npy'))
    want_all = np.concatenate([full.exact_sums_packed(k, case_x(name)[k]) for k in range(2)])
    np.testing.assert_allclose(batch, want_all, rtol=1e-12, atol=1e-13)
    sliced = np.load(os.path.join(tmp_path, 'sliced1.npy'))
    np.testing.assert_allclose(sliced, want_all, rtol=1e-12, atol=1e-13)
    wide = np.load(os.path.join(tmp_path, 'wide1.npy'))
    full2 = OracleEngine()
    full2.bind(types.SimpleNamespace(list_times=times, T_f=T_f), CASES[WIDE_CASE][0])
    want_wide = np.concatenate([full2.exact_sums_packed(k, _wide_x(WIDE_CASE)[k]) for k in range(2)])
    np.testing.assert_allclose(wide, want_wide, rtol=1e-12, atol=1e-13)
    held = [np.load(os.path.join(tmp_path, 'held%d.npy' % r)) for r in range(2)]
    # rank 1 holds its half plus a halo, not the whole path
    assert np.all(held[1] < np.array([700, 500])) and np.all(held[0] < np.array([700, 500]))
    r0 = np.load(os.path.join(tmp_path, 'ranges0.npy'))
    r1 = np.load(os.path.join(tmp_path, 'ranges1.npy'))
    assert np.all(r0[:, 1] == r1[:, 0])


def _episodes_worker(rank, world, port, paths, T_f, name, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here, os.path.join(here, 'golden')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import aslsd_b200
    from aslsd_b200.events import Episodes, ProcessPath
    from oracle_engine import OracleEngine
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    x, log = _fit_episodes(paths, T_f, name, dist.group.WORLD)
    np.save(os.path.join(out_dir, 'x%d.npy' % rank), np.concatenate(x))
    dist.barrier()
    dist.destroy_process_group()


def _fit_episodes(paths, T_f, name, group):
    import aslsd_b200
    from aslsd_b200.events import Episodes, ProcessPath
    from oracle_engine import OracleEngine
    model = CASES[name][0]
    d = model['d']
    objs = []
    for times in paths:
        eng = OracleEngine()
        objs.append(ProcessPath(times, T_f, engine=eng))
    ests = [[aslsd_b200.MHPExactEstimB200(engine=p.engine) for p in objs] for _ in range(d)]
    mirror = aslsd_b200.MHP(model)
    return mirror.fit_episodes(Episodes(objs, T_f=T_f), [np.asarray(v, float) for v in case_x(name)],
                               n_iter=4, estimators=ests, process_group=group)


@pytest.mark.timeout(300)
def test_fit_episodes_sharded_over_two_ranks(tmp_path):
    """Episodes dealt to 2 ranks, per-path gradients all-reduced: same iterates as one process."""
    name = 'exp2d'
    rng = np.random.default_rng(21)
    T_f = 80.
    paths = [[np.sort(rng.uniform(0, T_f, size=60 + 7 * q)), np.sort(rng.uniform(0, T_f, size=45 + 5 * q))]
             for q in range(3)]
    port = _free_port()
    mp.spawn(_episodes_worker, args=(2, port, paths, T_f, name, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(os.path.join(tmp_path, 'x%d.npy' % r)) for r in range(2)]
    np.testing.assert_array_equal(got[0], got[1])
    x, _ = _fit_episodes(paths, T_f, name, None)
    np.testing.assert_allclose(got[0], np.concatenate(x), rtol=1e-12)
