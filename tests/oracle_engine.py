"""Test double of ``aslsd_b200.engine.Engine`` backed by the numpy oracle.

TEST INFRASTRUCTURE ONLY: it lets the CPU suite exercise the host-side logic
of the estimators (sampling order, allocation bookkeeping, result shapes,
logs) against the live reference, where no GPU exists.  The product never
imports it; on the GPU box the same tests run against the real CUDA engine.
"""
import numpy as np

from aslsd_b200.engine import term_layout
from aslsd_b200.model import describe_model
from oracle import exact as OE
from oracle import stratified as OS


class OracleEngine:
    def __init__(self):
        self.desc = None
        self.times = None
        self.T_f = None
        self.n_calls = 0

    def bind(self, process_path, model):
        self.times = [np.asarray(t, dtype=float) for t in process_path.list_times]
        self.T_f = float(process_path.T_f)
        self.desc = describe_model(model)
        self.plain = self.desc.to_plain()
        return self.desc

    def exact_sums_packed(self, k, x_k, compute_f_sum=True, compute_diff_f_sum=True,
                          ranges=None, count_pairs=False):
        self.n_calls += 1
        sum_f, sum_diff_f, _ = OE.exact_sums_windowed(self.plain, k, x_k, self.times, self.T_f,
                                                      ranges=ranges)
        layout, out_len = term_layout(self.desc, k)
        out = np.zeros(out_len)
        for name, off, n1, n2, n_func in layout:
            out[off] = sum_f[name]
            if not compute_diff_f_sum:
                continue
            if n_func == 2:
                out[off + 1:off + 1 + n1] = sum_diff_f[name][1]
                out[off + 1 + n1:off + 1 + n1 + n2] = sum_diff_f[name][2]
            else:
                out[off + 1:off + 1 + n1] = sum_diff_f[name]
        return out

    # split form used by aslsd_b200.distributed: for the test double the "raw
    # accumulators" are the packed per-term sums themselves (additive over shards)
    def exact_raw_len(self, k):
        return term_layout(self.desc, k)[1]

    def exact_raw_used(self):
        return self._raw_used

    def exact_raw_device(self, k, x_k, raw_ptr, compute_f_sum=True, compute_diff_f_sum=True,
                         ranges=None):
        import ctypes
        n = self.exact_raw_len(k)
        buf = np.ctypeslib.as_array((ctypes.c_double * n).from_address(raw_ptr))
        buf[:] = self.exact_sums_packed(k, x_k, compute_f_sum, compute_diff_f_sum, ranges=ranges)
        self._raw_used = n

    def exact_assemble(self, k, x_k, raw_host, compute_f_sum=True, compute_diff_f_sum=True):
        return np.array(raw_host, dtype=float)

    def exact_raw_len_batch(self, n_k):
        return n_k * max(term_layout(self.desc, k)[1] for k in range(self.desc.d))

    def exact_raw_device_batch(self, ks, xs, raw_ptr, compute_f_sum=True, compute_diff_f_sum=True,
                               ranges=None):
        import ctypes
        n = self.exact_raw_len_batch(len(ks))
        buf = np.ctypeslib.as_array((ctypes.c_double * n).from_address(raw_ptr))
        buf[:] = 0.
        stride = n // len(ks)
        for q, (k, x) in enumerate(zip(ks, xs)):
            packed = self.exact_sums_packed(k, x, compute_f_sum, compute_diff_f_sum, ranges=ranges)
            buf[q * stride:q * stride + len(packed)] = packed
        self._raw_used = n

    def exact_assemble_batch(self, ks, xs, raw_host, compute_f_sum=True, compute_diff_f_sum=True):
        stride = len(raw_host) // len(ks)
        return [np.array(raw_host[q * stride:q * stride + term_layout(self.desc, k)[1]], dtype=float)
                for q, k in enumerate(ks)]

    def sampled_sums(self, k, x_k, terms, stratum_offsets, ix_m, ix_n=None, strata_sizes=None,
                     default_alloc=None, compute_diff_f_sum=True, grad_alloc=False):
        self.n_calls += 1
        sums, out_off, means, stds = [], [0], [], []
        s_ix = 0
        groups = []
        for kind, i, j, n_strata in terms:
            group = None
            for s in range(n_strata):
                a, b = stratum_offsets[s_ix], stratum_offsets[s_ix + 1]
                vals = OS.term_values(self.plain, k, x_k, self.times, self.T_f, kind, i, j,
                                      ix_m[a:b], None if ix_n is None else ix_n[a:b])
                if not compute_diff_f_sum:
                    vals[:, 1:] = 0.
                tot, mean, std = OS.stratum_statistics(vals, grad_alloc)
                sums.append(tot)
                out_off.append(out_off[-1] + len(tot))
                means.append(mean)
                stds.append(std)
                if strata_sizes is not None and strata_sizes[s_ix] > 0.:
                    if group is None:
                        group = [s_ix, 0]
                        groups.append(group)
                    group[1] += 1
                else:
                    group = None
                s_ix += 1
        rel = None
        if strata_sizes is not None:
            rel = np.array(default_alloc, dtype=float)
            for a, n in groups:
                rel[a:a + n] = OS.update_allocation(strata_sizes[a:a + n], stds[a:a + n],
                                                    default_alloc[a:a + n])
        return (np.concatenate(sums) if sums else np.zeros(0), np.asarray(out_off),
                np.asarray(means), np.asarray(stds), rel)
