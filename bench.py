"""Benchmark of the exact LSE gradient path (BASELINE.json metric:
"exact LSE-gradient event-pair terms/sec at 1/2/4/8 B200; s/epoch ASLSD fit").

Workload: config C5 of SURVEY.md section 8d - a 10-dimensional MHP whose 100
kernels are sums of 10 Gaussians (beta, delta fixed; omega free), ~10^7
simulated events.  One step = one full exact evaluation (function value and
all parameter derivatives of every term) for all d sub-problems k, i.e. what d
calls of ``lse_k_grad_estimate`` compute, issued as ONE batched call
(aslsd_exact_sums_batch); nothing is cached between steps.  Within a step the
raw basis-pair sums that several sub-problems share (beta / delta are fixed in
this model, so upsilon_i_j rows do not depend on k) are evaluated once, and the
interior targets of every double sum go through the lag quadrature (one pass
over the event pairs of each ordered pair of event streams, csrc/lag_quadrature.cuh);
the work actually executed is reported in ``work``.
The unit counted is the in-window event pair of a named double sum (phi_i,
upsilon_i_j) of one k.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU oracle arm, same stream

N > 1 is launched under torchrun (one rank per GPU); the events are replicated,
target events are sharded by time block, raw accumulators are combined with one
NCCL allreduce per step (strong scaling: total work is fixed).

Parity gate: before anything is timed, rank 0 evaluates a time slice of
sub-problem 0 with the numpy oracle and compares every per-term sum of the device
(single call and the batched call the benchmark times) at 1e-10; the run exits
non-zero if the gate fails.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# flops per unit of executed work, SURVEY 8d conventions (add/mul 1, FMA 2, div/sqrt 10,
# EXP 25, NDTR 60).  Per-item pair kernels and quadrature rows: interior Gaussian upsilon
# value of one basis pair = 1 EXP + 1 NDTR + 12.  Lag-moment kernel, algorithmic flops per
# event pair: lag (1 add), bin and position in the bin (1 FMA, 3 add), doubling (2 add),
# three-term recurrence (LQ_ND - 2 FMA), LQ_ND - 1 accumulations = 37 at LQ_ND = 12.
LQ_ND = 12
FLOPS_UPS_EVAL = 25 + 60 + 12
FLOPS_SINGLE_EVAL = 60 + 10
FLOPS_MOMENT_PAIR = 1 + 5 + 2 + 2 * (LQ_ND - 2) + (LQ_ND - 1)
# the reference-equivalent figures of SURVEY 8d for one SBF-10 pair-term
FLOPS_REF_DIRECT = 15700     # every basis pair evaluated per event pair, as the reference does
FLOPS_REF_SHARED = 3200      # with (beta, delta)-grid sharing
PARITY_GATE = 1e-10
# dram bytes of the dominant kernel: written by scripts/ncu_traffic.py from the committed ncu capture
TRAFFIC_FILE = os.path.join(ROOT, 'profiles', 'lag_moment_kernel_dram.json')


def executed_flops(work):
    """fp64 flops actually executed by the kernels of one call, from the engine's own
    counters (aslsd_last_work_counts)."""
    return ((work['pair_evals'] + work['quad_evals']) * FLOPS_UPS_EVAL
            + work['single_evals'] * FLOPS_SINGLE_EVAL + work['moment_pairs'] * FLOPS_MOMENT_PAIR)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--T_f', type=float, default=1.1e6, help='observation window (1.1e6 -> ~10^7 events)')
    ap.add_argument('--d', type=int, default=10)
    ap.add_argument('--n_basis', type=int, default=10)
    ap.add_argument('--cpu_sample_T', type=float, default=600., help='time slice of the CPU baseline sample')
    ap.add_argument('--no_cpu_baseline', action='store_true',
                    help='skip the CPU legs (oracle timing, parity gate, ASLSD epochs, plug-in e2e)')
    ap.add_argument('--per_k', action='store_true', help='one call per sub-problem (no sharing across k)')
    ap.add_argument('--no_quadrature', action='store_true', help='per-item pair kernels only')
    return ap.parse_args()


class ClockSampler:
    """SM clocks / throttle reasons during the timed region, polled through NVML every
    few milliseconds from a thread (a step lasts ~10 ms: nvidia-smi's loop is too coarse)."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown',
               0x4: 'sw_power_cap'}

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.sm, self.mask = [], 0
        self.handle = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.handle = None

    def sample(self):
        if self.handle is None:
            return
        try:
            self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
            self.mask |= int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        except Exception:
            pass

    def _loop(self):
        while not self.stop_flag.is_set():
            self.sample()
            time.sleep(0.002)

    def start(self):
        if self.handle is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.handle is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml unavailable'], 'samples': 0}
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=1)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.sm_max,
                'reasons': reasons, 'samples': len(self.sm)}


def load_measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0}, 'fallback'


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one lag_moment_warp_kernel launch on this
    workload, from the committed ``ncu --set full`` capture (profiles/); None when there is none."""
    try:
        with open(TRAFFIC_FILE) as f:
            t = json.load(f)
        return float(t['dram_bytes_read']) + float(t['dram_bytes_write']), t
    except Exception:
        return None, None


def make_workload(args):
    """Both arms: the same simulated stream (seed 1234) and the same evaluation points."""
    from aslsd_b200 import simulate
    return simulate.config_c5(d=args.d, n_basis=args.n_basis, T_f=args.T_f, seed=1234)


def workload_config(args, n_events):
    """``config`` of the JSON line: identical in both arms."""
    return {'workload': 'C5: %d-D SBF-Gaussian(%d) MHP, %d simulated events, exact LSE value+gradient sums of '
                        'all %d sub-problems per step' % (args.d, args.n_basis, n_events, args.d),
            'n_events': n_events, 'T_f': args.T_f, 'd': args.d, 'n_basis': args.n_basis, 'seed': 1234,
            'evaluation_point': 'truth x (1 + 0.1 U(-1, 1)), default_rng(12)'}


def slice_ranges(times, sample_T):
    return [(0, int(np.searchsorted(t, sample_T))) for t in times]


def _oracle_job(model, k, x, times, T_f, ranges):
    from oracle import exact as OE
    return OE.exact_sums_windowed(model, k, x, times, T_f, ranges=ranges)


def cpu_oracle(model, times, T_f, xs, sample_T, n_proc=1):
    """The numpy oracle (windowed restatement of the reference's numpy path, oracle/exact.py) on
    a bounded sample of the workload: targets inside the first ``sample_T`` time units,
    sub-problem k = 0, value + all derivatives.  Returns (pair-terms/s, seconds, pair-terms,
    (sum_f, sum_diff_f) when run in one piece)."""
    # only the events the slice can reach are handed to the workers
    t_keep = sample_T + 64.
    short = [t[:int(np.searchsorted(t, t_keep))] for t in times]
    if n_proc <= 1:
        t0 = time.perf_counter()
        f, df, n_pairs = _oracle_job(model, 0, xs[0], short, T_f, slice_ranges(short, sample_T))
        dt = time.perf_counter() - t0
        return sum(n_pairs.values()) / dt, dt, sum(n_pairs.values()), (f, df)
    import multiprocessing as mp
    cuts = np.linspace(0., sample_T, n_proc + 1)
    jobs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        r = [(int(np.searchsorted(t, a)), int(np.searchsorted(t, b))) for t in short]
        jobs.append((model, 0, xs[0], short, T_f, r))
    t0 = time.perf_counter()
    with mp.get_context('fork').Pool(n_proc) as pool:
        res = pool.starmap(_oracle_job, jobs)
    dt = time.perf_counter() - t0
    tot = sum(sum(r[2].values()) for r in res)
    return tot / dt, dt, tot, None


def pack_oracle(desc, k, sum_f, sum_diff_f):
    """Oracle dictionaries -> the packed layout of aslsd_exact_sums."""
    from aslsd_b200.engine import term_layout
    layout, n = term_layout(desc, k)
    out = np.zeros(n)
    for name, off, n1, n2, n_func in layout:
        out[off] = sum_f[name]
        dv = sum_diff_f[name]
        if n_func == 2:
            out[off + 1:off + 1 + n1] = dv[1]
            out[off + 1 + n1:off + 1 + n1 + n2] = dv[2]
        else:
            out[off + 1:off + 1 + n1] = dv
    return out


def parity_gate(engine, model, times, T_f, xs, sample_T, oracle_sums):
    """Device against the oracle on the CPU leg's slice: per-term sums, rel <= 1e-10 with the
    absolute floor 1e-300 (SURVEY 8d "Parity gates in the bench")."""
    ranges = slice_ranges(times, sample_T)
    want = pack_oracle(engine.desc, 0, *oracle_sums)
    single = engine.exact_sums_packed(0, xs[0], ranges=ranges)
    batch = engine.exact_sums_batch(list(range(len(xs))), xs, ranges=ranges)[0]
    rel = 0.
    for got in (single, batch):
        rel = max(rel, float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))))
    return {'max_rel': rel, 'gate': PARITY_GATE, 'ok': bool(rel <= PARITY_GATE), 'n_sums': int(len(want)),
            'what': 'every per-term sum and derivative of sub-problem 0 over the targets of the first %g time '
                    'units, single call and batched call, against oracle/exact.py' % sample_T}


ASLSD_CONFIGS = ('c2', 'c3', 'c4')


def aslsd_configs():
    from aslsd_b200 import simulate
    exp = lambda: [{'type': 'exp'}]
    return {
        # C2: 2-D exponential (fit_exp2d.py:34-51), ~10^5 events
        'c2': ({'d': 2, 'kernels': [[exp(), exp()], [exp(), exp()]]}, [1.5, 1.],
               [[[.2, 1.], [.6, 1.5]], [[.7, 2.], [.1, 1.3]]], 7.7e3),
        # C3: 1-D Gaussian (fit_gauss1d.py:31-34), ~10^6 events
        'c3': ({'d': 1, 'kernels': [[[{'type': 'gauss'}]]]}, [1.5], [[[.5, .5, 3.]]], 3.33e5),
        # C4: 1-D SBF-Gaussian(10), ~10^6 events (fit_sbgauss1d10r.py:60-63)
        'c4': (simulate.sbf_gauss_model(1, 10), [1.0],
               [[simulate.random_mixture_weights(10, coeff=0.5)]], 5e5),
    }


def aslsd_epochs(engine):
    """Second part of BASELINE.json's metric: seconds per epoch of an ASLSD fit (one stratified
    gradient estimate + ADAM step for every k), configs C2-C4 of SURVEY 8d."""
    import aslsd_b200
    from aslsd_b200 import simulate
    out = {}
    for name, (model, mu, kp, T_f) in aslsd_configs().items():
        d = model['d']
        times = simulate.simulate_mhp(model, mu, kp, T_f, seed=1234)
        path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
        mirror = aslsd_b200.MHP(model)
        rng = np.random.default_rng(12)
        x0 = []
        for k in range(d):
            x = np.concatenate([[mu[k]]] + [np.asarray(kp[k][i], float) for i in range(d)])
            x0.append(x * (1 + 0.1 * rng.uniform(-1, 1, size=len(x))))
        mirror.fit(path, [x.copy() for x in x0], n_iter=3, engine=engine, seed=1)
        n_iter = 20
        t0 = time.perf_counter()
        mirror.fit(path, [x.copy() for x in x0], n_iter=n_iter, engine=engine, seed=1)
        out[name] = {'s_per_epoch': (time.perf_counter() - t0) / n_iter,
                     'n_events': int(sum(len(t) for t in times))}
    return out


def plugin_e2e(engine, model, times, T_f, xs, n_rep):
    """One epoch through the reference's plug-in call: a fresh upload of the events (host numpy
    arrays, pageable, as a reference caller holds them), then d x
    ``MHPExactEstimB200.lse_k_grad_estimate(x_k)`` as mhp.py:1778-1787 issues them, every gradient
    returned as a host array.  Weight cache off (every call recomputes the quadrature) and on (the
    library default: the moments of the event pairs are kept while events and bin grids stay)."""
    import aslsd_b200
    d = model['d']
    path = aslsd_b200.ProcessPath(times, T_f, book_keeping=False, engine=engine)
    path.list_times2end = None       # O(N) host arrays the exact estimator never reads
    mirror = aslsd_b200.MHP(model)
    ests = []
    for k in range(d):
        est = aslsd_b200.MHPExactEstimB200(engine=engine, reference_recombine=False)
        est.initialize(k, 4, mirror, path)
        ests.append(est)
    out = {}
    for label, cache in (('cache_off', False), ('cache_on', True)):
        engine.use_weight_cache(cache)
        per_call, upload, total = [], [], []
        # two untimed epochs: the first call on a row plan assembles on the host, the second records the
        # plan's assembly map for the device (one-off, tens of ms per plan); a fit runs hundreds of epochs
        for rep in range(n_rep + 2):
            engine.synchronize()
            t0 = time.perf_counter()
            engine._events_key = None                         # a new fit's first bind: one upload of the events
            engine.bind(path, mirror)
            t1 = time.perf_counter()
            calls = []
            for k in range(d):
                tc = time.perf_counter()
                g = ests[k].lse_k_grad_estimate(xs[k])
                calls.append(time.perf_counter() - tc)
            assert np.all(np.isfinite(g))
            t2 = time.perf_counter()
            if rep > 1:
                upload.append(t1 - t0)
                per_call.append(calls)
                total.append(t2 - t0)
        pc = np.mean(np.array(per_call), axis=0)
        out[label] = {'ms_per_epoch': 1e3 * float(np.mean(total)), 'upload_ms': 1e3 * float(np.mean(upload)),
                      'first_call_ms': 1e3 * float(pc[0]), 'later_calls_ms_mean': 1e3 * float(np.mean(pc[1:]))}
        # the following epochs of the same fit: the events stay (the reference's fit loop keeps one path),
        # the evaluation points move
        steady = []
        for rep in range(3):
            engine.synchronize()
            t0 = time.perf_counter()
            for k in range(d):
                g = ests[k].lse_k_grad_estimate(xs[k] * (1. + 1e-3 * (rep + 1)))
            steady.append(time.perf_counter() - t0)
        out[label]['next_epochs_ms'] = 1e3 * float(np.mean(steady[1:]))
    engine.use_weight_cache(False)
    return out


def device_simulation(engine, args):
    """The benchmark's model simulated on the device (aslsd_simulate, the generation loop of
    MHP.simulate): reported beside the headline, not part of the timed step - both arms take their
    input from the numpy simulator so that the reference arm has no device code on its path."""
    from aslsd_b200 import simulate
    model = simulate.sbf_gauss_model(args.d, args.n_basis)
    w = simulate.random_mixture_weights(args.n_basis, seed=1234)
    kp = [[0.5 * w / args.d for _ in range(args.d)] for _ in range(args.d)]
    mu = np.full(args.d, 0.5)
    ms, n = [], 0
    for rep in range(3):
        t0 = time.perf_counter()
        out = engine.simulate(model, mu, kp, args.T_f, seed=1234 + rep)
        ms.append(1e3 * (time.perf_counter() - t0))
        n = int(sum(len(t) for t in out))
    engine.close()
    return {'ms': float(np.min(ms)), 'n_events': n, 'events_per_s': n / (1e-3 * float(np.min(ms))),
            'note': 'cluster simulation of the C5 model on the device, event times back on the host (wall clock, '
                    'best of 3); the numpy simulator that feeds both arms takes seconds'}


def run_reference(args):
    """The reference's CPU path (pure Python / numpy: timed as the oracle port) on a bounded
    sample of the same stream and evaluation point as the GPU arm, all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    model, times, T_f, xs = make_workload(args)
    n_events = int(sum(len(t) for t in times))
    sample_T = args.cpu_sample_T
    cores = os.cpu_count() or 1
    n_proc = max(1, min(cores, 32))
    rates, times_s = [], []
    pairs = 0
    for step in range(args.warmup + args.steps):
        rate, dt, pairs, _ = cpu_oracle(model, times, T_f, xs, sample_T, n_proc=n_proc)
        if step >= args.warmup:
            rates.append(rate)
            times_s.append(dt)
    value = float(np.mean(rates))
    sample = ('sub-problem k=0, targets in the first %g time units of the stream (%d in-window pair-terms per '
              'step), split over %d processes' % (sample_T, pairs, n_proc))
    line = {
        'impl': 'reference', 'metric': 'exact LSE-gradient event-pair terms/sec',
        'value': value, 'unit': 'pair-terms/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * float(np.mean(times_s)),
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': workload_config(args, n_events),
        'sample': sample,
        'note': 'reference is pure Python/numpy: timed as the oracle port (oracle/exact.py windowed '
                'restatement of its numpy path); each step is a bounded sample of the workload',
        'cpu_baseline': {'value': value, 'unit': 'pair-terms/s', 'cores': n_proc, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'pair-terms/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    import aslsd_b200
    from aslsd_b200.distributed import ShardedExactSums

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)

    model, times_pageable, T_f, xs = make_workload(args)
    times_pageable = [np.ascontiguousarray(t, dtype=np.float64) for t in times_pageable]
    # the step's inputs start in page-locked host memory (the library uploads such arrays from
    # where they are; pageable arrays go through its own pinned staging block instead)
    times = [torch.from_numpy(t).pin_memory().numpy() for t in times_pageable]
    d = model['d']
    n_events = int(sum(len(t) for t in times))
    engine = aslsd_b200.Engine(local_rank)
    engine.set_stream(torch.cuda.current_stream().cuda_stream)
    engine.set_events(times, T_f)
    engine.set_model(model)
    engine.use_lag_quadrature(not args.no_quadrature)
    engine.use_weight_cache(False)     # every step recomputes the quadrature from the event pairs
    engine.use_async_events(True)      # page-locked inputs: planning overlaps the transfer
    ks = list(range(d))

    # ---- CPU leg + parity gate (rank 0, before anything is timed) ---------------------------------
    cpu_baseline = parity = None
    if rank == 0 and not args.no_cpu_baseline:
        rate, dt, pairs, oracle_sums = cpu_oracle(model, times, T_f, xs, args.cpu_sample_T, n_proc=1)
        cpu_baseline = {
            'value': rate, 'unit': 'pair-terms/s', 'cores': 1, 'kind': 'port',
            'sample': 'numpy oracle (oracle/exact.py), sub-problem k=0, targets in the first %g time '
                      'units (%d pair-terms, %.1f s); host has %d cores' %
                      (args.cpu_sample_T, pairs, dt, os.cpu_count() or 0)}
        parity = parity_gate(engine, model, times, T_f, xs, args.cpu_sample_T, oracle_sums)

    sharded = ShardedExactSums(engine, times, rank=rank, world_size=world,
                               dist=dist if world > 1 else None, device=device)

    # work accounting (outside the timed region): in-window pair-terms of the
    # whole job and closed-form evaluations of this rank
    pair_terms = 0
    if rank == 0:
        for k in range(d):
            _, counts = engine.exact_sums_packed(k, xs[k], count_pairs=True)
            pair_terms += sum(counts.values())
    fp64_peak = engine.measure_fp64_peak()
    idx_ms, idx_bytes = (None, None)
    if rank == 0:
        engine.index_maps_kernel_time()
        idx_ms, idx_bytes = engine.index_maps_kernel_time()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    def step():
        if args.per_k:
            return [sharded.launch(k, xs[k]) for k in range(d)]
        return [sharded.launch_packed_batch(ks, xs)]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        step()
    sync_all()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    sync_all()
    for s in range(args.steps):
        flush.zero_()
        ev[s][0].record()
        step()
        ev[s][1].record()
    if rank == 0:
        sampler.sample()     # at least one sample while the last steps are in flight
    sync_all()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    t_max = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    total_ms = float(t_max.item())
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel accounting (library CUDA events on the launching stream): all kernels of a
    # step, and the lag-moment kernel alone, averaged over a few L2-flushed steps
    k_ms_sum, m_ms_sum, n_launch_step = 0., 0., 0
    work = {}
    n_acc = max(1, min(args.steps, 3))
    for rep in range(n_acc):
        flush.zero_()
        n_launch_step = 0
        work = {}
        for call in ([(ks, xs)] if not args.per_k else [([k], [xs[k]]) for k in range(d)]):
            sharded.launch_packed_batch(*call)
            ms, nl = engine.last_timing()
            k_ms_sum += ms / n_acc
            m_ms_sum += engine.last_moment_timing() / n_acc
            n_launch_step += nl
            for key, val in engine.last_work_counts().items():
                work[key] = work.get(key, 0) + val
    torch.cuda.synchronize()

    # the iterations of a fit: same events, same (beta, delta) -> the quadrature weights are
    # reused and only the node evaluations run (reported beside the headline, not as it)
    engine.use_weight_cache(True)
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_fit = 5
    for _ in range(n_fit):
        step()
    torch.cuda.synchronize()
    fit_iter_ms = 1e3 * (time.perf_counter() - t0) / n_fit
    engine.use_weight_cache(False)

    # end-to-end through the public API with host buffers: upload the events,
    # evaluate every k, read the packed sums back.  Page-locked inputs (the contract's case),
    # then pageable ones (what a numpy caller of the reference holds).
    def e2e_pass(host_times):
        out_ms = []
        d2h = 0
        for s in range(1 + max(1, args.steps)):      # first pass untimed (buffer growth)
            sync_all()
            t0 = time.perf_counter()
            if world > 1 and not args.per_k:
                sharded.set_events(host_times, T_f, ks, xs)     # this rank's time block + max-lag halo only
            else:
                engine.set_events(host_times, T_f)
            if args.per_k:
                d2h = 0
                for k in range(d):
                    sharded.packed(k, xs[k])
                    d2h += 8 * engine.exact_raw_used()
            else:
                sharded.packed_batch(ks, xs)
                d2h = sharded.last_d2h_bytes
            sync_all()
            if s > 0:
                out_ms.append(1e3 * (time.perf_counter() - t0))
        t = torch.tensor([float(np.mean(out_ms))], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), d2h

    h2d = sum(8 * len(x) for x in xs)
    e2e_ms, d2h = e2e_pass(times)
    h2d_events = torch.tensor([8. * float(sum(engine.n_events))], dtype=torch.float64, device=device)
    d2h_all = torch.tensor([float(d2h)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(h2d_events, op=dist.ReduceOp.SUM)
        dist.all_reduce(d2h_all, op=dist.ReduceOp.SUM)
    engine.use_async_events(False)
    e2e_pageable_ms, _ = e2e_pass(times_pageable)
    engine.use_async_events(True)

    plugin = aslsd = simulation = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.per_k:
        engine.set_events(times, T_f)
        plugin = plugin_e2e(engine, model, times_pageable, T_f, xs, n_rep=2)
        # stratified (ASLSD) fit epochs on their own engine (the C5 events stay untouched)
        aslsd = aslsd_epochs(aslsd_b200.Engine(local_rank))
        simulation = device_simulation(aslsd_b200.Engine(local_rank), args)

    if rank == 0:
        peaks, peak_kind = load_measured_peaks()
        ms_per_step = total_ms / args.steps
        value = pair_terms / (ms_per_step * 1e-3)
        # roofline of the dominant kernel (lag_moment_warp_kernel) on this rank: its algorithmic
        # flops (event pairs it binned x FLOPS_MOMENT_PAIR) over its own CUDA-event time
        flops = executed_flops(work)
        m_flops = work.get('moment_pairs', 0) * FLOPS_MOMENT_PAIR
        if m_ms_sum > 0:
            achieved = m_flops / (m_ms_sum * 1e-3) * 1e-12
        else:
            achieved = flops / (k_ms_sum * 1e-3) * 1e-12 if k_ms_sum > 0 else None
        n_src_tgt = 2 * 8 * n_events * d / world   # every stream reads its targets and sources once
        traffic, traffic_src = load_traffic()
        if world != 1 or args.no_quadrature or args.per_k:
            traffic, traffic_src = None, None     # the capture is of the default single-GPU step
        config = workload_config(args, n_events)
        config.update({
            'pair_terms_per_step': pair_terms,
            'event_pairs_binned_per_step': work.get('moment_pairs', 0),
            'unit_note': 'pair-terms count every upsilon_i_j / phi_i pair once per sub-problem k (SURVEY 8d); '
                         'beta and delta are fixed in this model, so the rows of the d sub-problems are shared and '
                         'the kernels bin each physical event pair once per step: event_pairs_binned_per_step',
            'l2': 'flushed between steps (256 MB write)',
            'call': 'one call per k' if args.per_k else 'one batched call for all k',
            'sharding': 'time blocks of target events; value: events replicated on every GPU; '
                        'e2e: every rank uploads its block + max-lag halo only; 1 NCCL allreduce '
                        'of the packed sums per step' if world > 1 else 'single GPU',
            'fit_iteration_ms': fit_iter_ms,
        })
        if aslsd:
            for name in ASLSD_CONFIGS:
                config['aslsd_%s_s_per_epoch' % name] = aslsd[name]['s_per_epoch']
                config['aslsd_%s_n_events' % name] = aslsd[name]['n_events']
        line = {
            'metric': 'exact LSE-gradient event-pair terms/sec',
            'value': value, 'unit': 'pair-terms/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config,
            'work': dict(work, note='rank-0 counters of one step: rows = unique work items '
                         'evaluated, items = work items of the d sub-problems before sharing; '
                         'moment_pairs = event pairs binned by the lag-moment kernel (streams, '
                         'bins: its quadrature streams / bins), quad_evals = closed-form '
                         'evaluations at quadrature nodes, pair_evals = per-item kernel '
                         'basis-pair evaluations (targets near T_f), single_evals = psi / '
                         'upsilonzero evaluations'),
            'index_maps': {'kernel_ms': idx_ms, 'algorithmic_bytes': idx_bytes,
                           'GBps': idx_bytes / (idx_ms * 1e-3) * 1e-9 if idx_ms else None,
                           'frac_of_hbm_peak': (idx_bytes / (idx_ms * 1e-3) * 1e-9 / peaks.get('hbm_gbs'))
                           if idx_ms else None,
                           'note': 'kappa / varpi / lag_sizes of all d^2 ordered pairs in one launch '
                                   '(HBM-bound batched searchsorted), tables left on the device'},
            'fit_iteration': {'ms': fit_iter_ms,
                              'note': 'same step with the quadrature weights kept between calls '
                                      '(ASLSD_OPT_WEIGHT_CACHE, the default outside this benchmark): '
                                      'what one more iteration of a fit at fixed (beta, delta) costs; '
                                      'not the headline'},
            'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                         'frac': (achieved / fp64_peak) if achieved else None,
                         'traffic': traffic,
                         'traffic_source': traffic_src,
                         'traffic_note': 'dram__bytes_read.sum + dram__bytes_write.sum of one lag_moment_warp_kernel '
                                         'launch, read from %s (written by scripts/ncu_traffic.py from the committed '
                                         'ncu --set full capture of this workload); algorithmic bytes %.0f MB (each '
                                         'event time read once per stream it belongs to) - the 80 MB of events stay '
                                         'in L2' % (os.path.relpath(TRAFFIC_FILE, ROOT), n_src_tgt / 1e6),
                         'kernel': 'lag_moment_warp_kernel',
                         'kernel_ms': m_ms_sum, 'kernel_share_of_step': (m_ms_sum / k_ms_sum) if k_ms_sum else None,
                         'algorithmic_flops_per_launch': m_flops,
                         'note': 'the kernel is bound by instruction issue (routing every event pair to the lane that '
                                 'owns its bin: ballots, shuffles, shared-memory stacks), not by the fp64 pipe or HBM '
                                 '(ncu summaries under profiles/). frac is its fp64 rate over the measured DFMA peak',
                         'all_kernels_tflops': flops / (k_ms_sum * 1e-3) * 1e-12 if k_ms_sum > 0 else None,
                         'reference_equivalent_tflops': pair_terms * FLOPS_REF_DIRECT / (ms_per_step * 1e-3) * 1e-12,
                         'reference_equivalent_shared_tflops': pair_terms * FLOPS_REF_SHARED / (ms_per_step * 1e-3) * 1e-12,
                         'reference_equivalent_note': 'pair-terms/s x 15 700 flops (SURVEY 8d figure for one SBF-10 '
                                                      'upsilon pair-term evaluated directly, as the reference does) '
                                                      'resp. x 3 200 (with grid sharing); algorithmic equivalents, '
                                                      'not hardware rates and not part of the roofline',
                         'peak_source': 'DFMA microbenchmark measured live (aslsd_measure_fp64_peak); '
                                        'MEASURED_PEAKS.json has no fp64 entry',
                         'executed_flops_per_step_rank0': flops, 'kernel_ms_per_step_rank0': k_ms_sum,
                         'hbm_peak_gbs': peaks.get('hbm_gbs'), 'hbm_peak_kind': peak_kind},
            'e2e': {'value': pair_terms / (e2e_ms * 1e-3), 'unit': 'pair-terms/s',
                    'h2d_bytes_per_step': int(h2d_events.item()) + h2d * world,
                    'd2h_bytes_per_step': int(d2h_all.item()),
                    'ms_per_step': e2e_ms,
                    'inputs': 'event times in page-locked host memory, uploaded in place every step, order check '
                              'on the device behind the transfer (ASLSD_OPT_ASYNC_EVENTS); one batched C-ABI call; '
                              'packed sums read back',
                    'pageable_ms_per_step': e2e_pageable_ms,
                    'pageable_value': pair_terms / (e2e_pageable_ms * 1e-3),
                    'pageable_note': 'the same step from pageable numpy arrays (what a caller of the reference '
                                     'holds): staged through the library\'s page-locked block, order check on the host'},
            'gpu_launches': n_launch_step * args.steps,
            'clocks': clocks,
            'parity': parity,
        }
        if plugin:
            line['e2e_plugin'] = dict(
                plugin, unit='ms per epoch (d estimator calls)',
                value_cache_off=pair_terms / (plugin['cache_off']['ms_per_epoch'] * 1e-3),
                value_cache_on=pair_terms / (plugin['cache_on']['ms_per_epoch'] * 1e-3),
                note='upload of the events from pageable numpy arrays + d x MHPExactEstimB200.lse_k_grad_estimate '
                     '(one call per sub-problem, as the reference\'s fit loop issues them, mhp.py:1778-1787), every '
                     'gradient back on the host.  Each call needs the moments of all d^2 event-stream pairs: with the '
                     'weight cache off they are recomputed d times per epoch (no sharing between calls - the batched '
                     'call of the headline exists for that), with the cache on (library default) the first call '
                     'computes them and the others reuse them.  next_epochs_ms: the following epochs of the same fit - same events, '
                     'moved evaluation points (omega only: beta and delta are fixed in this model)')
        if simulation:
            line['simulation'] = simulation
        if aslsd:
            line['aslsd_fit'] = dict(
                {name: aslsd[name] for name in ASLSD_CONFIGS},
                note='s/epoch of an ASLSD fit (MHPStratEstimB200 + ADAM through the mirror fit loop, host '
                     'sampling with the reference\'s RNG stream, sampled-lag kernel on the device); '
                     'reported beside the headline metric')
        if cpu_baseline:
            line['cpu_baseline'] = cpu_baseline
        print(json.dumps(line))
    ok = parity is None or parity['ok']
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        sys.stderr.write('parity gate FAILED: max rel %.3e > %.0e\n' % (parity['max_rel'], PARITY_GATE))
        sys.exit(1)


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
