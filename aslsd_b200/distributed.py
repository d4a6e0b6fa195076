"""Time-block sharding of the exact LSE sums over the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  The observation
window [0, T_f] is cut into ``world_size`` contiguous time blocks holding equal
numbers of target events; rank g evaluates every term for the targets of its
block only.  Sources are taken from the halo [block start - max lag, block
end).  Two ways to hold the events:
  * ``ShardedExactSums(engine, list_times, ...)`` after ``engine.set_events(all events)``:
    replicated events (80 MB at 10^7 events), the halo is purely logical;
  * ``sharded.set_events(list_times, T_f, ks, xs)``: every rank uploads only its block plus
    the halo of ``engine.exact_max_lag(ks, xs)`` time units before it (1/world_size of the
    host -> device traffic); re-upload when the parameters widen the kernels' reach.
Neither needs an exchange of events.  All sums are additive over targets
(mhp_exact_estimator.py:633, :667, :703), hence one allreduce of the raw accumulator
vector combines the shards; the packed per-term sums are then assembled on every rank.
"""
import numpy as np

from aslsd_b200.engine import unpack_terms


def time_block_ranges(list_times, world_size):
    """ranges[g][i] = (m_lo, m_hi): target events of type i owned by rank g.
    Block boundaries are quantiles of the pooled event times, so that every
    rank gets the same number of targets (hence, for a stationary stream, the
    same number of in-window pairs)."""
    if world_size == 1:
        return [[(0, len(t)) for t in list_times]]
    n_tot = sum(len(t) for t in list_times)
    # Quantiles of the pooled times without materialising the pooled sort: a bracketing search on
    # the cut time, all cuts at once (one vectorised searchsorted per type and step), alternating
    # an interpolation step (event counts are close to linear in time) with a bisection step.
    # A cut is accepted once the number of events below it is within `tol` of its quantile:
    # any cut time partitions the events consistently, the count only balances the load.
    t_min = min(t[0] for t in list_times if len(t))
    t_max = max(t[-1] for t in list_times if len(t))
    targets = np.array([n_tot * g // world_size for g in range(1, world_size)], dtype=np.float64)
    tol = max(len(list_times), n_tot >> 16)
    lo = np.full(world_size - 1, t_min, dtype=np.float64)
    hi = np.full(world_size - 1, t_max, dtype=np.float64)
    c_lo = np.zeros(world_size - 1)
    c_hi = np.full(world_size - 1, float(n_tot))
    done = np.zeros(world_size - 1, dtype=bool)
    for it in range(128):
        if it % 2 == 0:
            frac = np.clip((targets - c_lo) / np.maximum(c_hi - c_lo, 1.), 0.02, 0.98)
        else:
            frac = 0.5
        mid = np.where(done, hi, lo + frac * (hi - lo))
        below = sum(np.searchsorted(t, mid, side='left') for t in list_times).astype(np.float64)
        hit = np.abs(below - targets) <= tol
        go_up = below < targets
        # an accepted cut collapses its bracket onto itself
        lo = np.where(done, lo, np.where(hit | go_up, mid, lo))
        hi = np.where(done, hi, np.where(hit | ~go_up, mid, hi))
        c_lo = np.where(done, c_lo, np.where(hit | go_up, below, c_lo))
        c_hi = np.where(done, c_hi, np.where(hit | ~go_up, below, c_hi))
        done |= hit
        if np.all(done | (hi - lo <= 1e-12 * max(abs(t_max), 1.))):
            break
    cuts = list(hi)
    edges = [[0] + [int(np.searchsorted(t, c, side='left')) for c in cuts] + [len(t)]
             for t in list_times]
    return [[(edges[i][g], edges[i][g + 1]) for i in range(len(list_times))]
            for g in range(world_size)]


def block_slices(list_times, block, max_lag):
    """Events a rank must hold for its target block ``block[i] = (lo, hi)``: per type the
    events from ``max_lag`` time units before the block's earliest target up to the block's end.
    Returns (local event arrays, target ranges inside them)."""
    starts = [np.asarray(t)[b[0]] for t, b in zip(list_times, block) if b[1] > b[0]]
    local, ranges = [], []
    for t, (lo, hi) in zip(list_times, block):
        t = np.asarray(t)
        first = int(np.searchsorted(t, min(starts) - max_lag, side='left')) if starts else lo
        first = min(first, lo)
        local.append(t[first:hi])
        ranges.append((lo - first, hi - first))
    return local, ranges


class ShardedExactSums:
    """Exact sums of sub-problem k over ``dist`` ranks (one Engine per rank)."""

    def __init__(self, engine, list_times, rank=0, world_size=1, dist=None, device=None):
        import torch
        self.torch = torch
        self.engine = engine
        self.rank, self.world_size, self.dist = rank, world_size, dist
        self.device = device
        self.ranges = time_block_ranges(list_times, world_size)[rank]
        self._raw = {}
        self.max_lag = None
        self.last_d2h_bytes = 0

    def set_events(self, list_times, T_f, ks, xs, slack=1.25):
        """Upload this rank's time block of ``list_times`` plus the halo its first targets can
        reach (``slack`` x the largest lag of the sub-problems ``ks`` at ``xs``) and point the
        target ranges at the block inside the uploaded slice.  Every later launch checks that
        the reach of its parameters still fits the halo and uploads a wider slice if it does
        not (``_check_reach``)."""
        block = self._blocks(list_times)[self.rank]
        key = self._reach_key(ks, xs)
        if getattr(self, '_reach_value', (None, None, None))[:2] == (key, id(self.engine.desc)):
            reach = self._reach_value[2]          # a property of the model and the parameters alone
        else:
            reach = self.engine.exact_max_lag(ks, xs)
            self._reach_value = (key, id(self.engine.desc), reach)
        self.max_lag = slack * reach
        self._reach_checked = key                 # these parameters fit: no second look at launch
        local, ranges = block_slices(list_times, block, self.max_lag)
        self.engine.set_events(local, T_f)
        self.ranges = ranges
        self._sliced = (list_times, T_f, slack)
        return ranges

    def _blocks(self, list_times):
        """Time blocks of ``list_times``, kept while the same arrays are handed in again."""
        key = (tuple(id(t) for t in list_times), tuple(len(t) for t in list_times), self.world_size)
        if getattr(self, '_blocks_key', None) != key:
            self._blocks_val = time_block_ranges(list_times, self.world_size)
            self._blocks_key = key
            self._blocks_ref = list(list_times)     # keeps the ids unique
        return self._blocks_val

    def _check_reach(self, ks, xs):
        """Halo mode only: a parameter update may have widened the kernels' reach (42 / beta,
        delta + 9 sigma) beyond the halo uploaded with the block - the first targets of every
        block but the first would then silently lose sources.  Re-slice and re-upload."""
        if self.max_lag is None:
            return
        key = self._reach_key(ks, xs)
        if key == getattr(self, '_reach_checked', None):
            return
        reach = self.engine.exact_max_lag(ks, xs)
        if reach <= self.max_lag:
            self._reach_checked = key
        if reach > self.max_lag:
            list_times, T_f, slack = self._sliced
            self.set_events(list_times, T_f, ks, xs, slack=max(slack, 1.25))

    @staticmethod
    def _reach_key(ks, xs):
        return (tuple(int(k) for k in ks), tuple(np.asarray(x, dtype=np.float64).tobytes() for x in xs))

    def raw_buffer(self, k):
        if k not in self._raw:
            n = self.engine.exact_raw_len(k)
            self._raw[k] = self.torch.zeros(n, dtype=self.torch.float64, device=self.device)
        return self._raw[k]

    def launch(self, k, x_k, compute_f_sum=True, compute_diff_f_sum=True):
        """Enqueue the shard's kernels and the allreduce; returns the device
        tensor holding the combined raw accumulators (asynchronous)."""
        raw = self.raw_buffer(k)
        if raw.is_cuda:
            # the library's kernels and the NCCL all-reduce share torch's current stream
            self.engine.set_stream(self.torch.cuda.current_stream().cuda_stream)
        self._check_reach([k], [x_k])
        self.engine.exact_raw_device(k, x_k, raw.data_ptr(), compute_f_sum,
                                     compute_diff_f_sum, ranges=self.ranges)
        raw = raw[:self.engine.exact_raw_used()]     # the rows this model has (same on all ranks)
        if self.world_size > 1:
            self.dist.all_reduce(raw, op=self.dist.ReduceOp.SUM)
        return raw

    # ---- all sub-problems of a step in one launch sequence + one all-reduce ---------------
    def launch_batch(self, ks, xs, compute_f_sum=True, compute_diff_f_sum=True):
        key = ('batch', len(ks))
        if key not in self._raw:
            n = self.engine.exact_raw_len_batch(len(ks))
            self._raw[key] = self.torch.zeros(n, dtype=self.torch.float64, device=self.device)
        raw = self._raw[key]
        if raw.is_cuda:
            self.engine.set_stream(self.torch.cuda.current_stream().cuda_stream)
        self._check_reach(ks, xs)
        self.engine.exact_raw_device_batch(ks, xs, raw.data_ptr(), compute_f_sum,
                                           compute_diff_f_sum, ranges=self.ranges)
        raw = raw[:self.engine.exact_raw_used()]     # the rows this model has (same on all ranks)
        if self.world_size > 1:
            self.dist.all_reduce(raw, op=self.dist.ReduceOp.SUM)
        return raw

    def launch_packed_batch(self, ks, xs, compute_f_sum=True, compute_diff_f_sum=True):
        """Kernels, device-side assembly and the all-reduce of the packed sums (assembly is linear:
        a few thousand doubles per sub-problem cross NVLink instead of the raw rows); returns the
        device tensor of the concatenated packed sums (asynchronous)."""
        key = ('packed', tuple(ks))
        if key not in self._raw:
            n = self.engine.exact_packed_len_batch(ks)
            self._raw[key] = self.torch.zeros(n, dtype=self.torch.float64, device=self.device)
        packed = self._raw[key]
        if packed.is_cuda:
            self.engine.set_stream(self.torch.cuda.current_stream().cuda_stream)
        self._check_reach(ks, xs)
        self.engine.exact_packed_device_batch(ks, xs, packed.data_ptr(), compute_f_sum,
                                              compute_diff_f_sum, ranges=self.ranges)
        if self.world_size > 1:
            self.dist.all_reduce(packed, op=self.dist.ReduceOp.SUM)
        return packed

    def packed_batch(self, ks, xs, compute_f_sum=True, compute_diff_f_sum=True):
        packed = self.launch_packed_batch(ks, xs, compute_f_sum, compute_diff_f_sum)
        host = packed.cpu().numpy()      # synchronises the stream
        self.last_d2h_bytes = host.nbytes
        return self.engine.split_packed(ks, host)

    def packed(self, k, x_k, compute_f_sum=True, compute_diff_f_sum=True):
        raw = self.launch(k, x_k, compute_f_sum, compute_diff_f_sum)
        host = raw.cpu().numpy()     # synchronises the stream
        return self.engine.exact_assemble(k, x_k, host, compute_f_sum, compute_diff_f_sum)

    def sums(self, k, x_k, compute_f_sum=True, compute_diff_f_sum=True):
        return unpack_terms(self.engine.desc, k,
                            self.packed(k, x_k, compute_f_sum, compute_diff_f_sum))
