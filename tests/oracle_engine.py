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

    def set_events(self, list_times, T_f):
        self.times = [np.asarray(t, dtype=float) for t in list_times]
        self.T_f = float(T_f)

    def set_horizon(self, T_f):
        self.T_f = float(T_f)

    def index_maps(self, want_kappa=True, want_varpi=True, want_lag_sizes=True):
        from oracle import time_ordering as OT
        kappa, varpi = OT.kappa_varpi_closed_form(self.times, self.T_f)
        return kappa, varpi, OT.lag_sizes(self.times, kappa, varpi)

    @property
    def d(self):
        return self.plain['d']

    @property
    def n_events(self):
        return np.array([len(t) for t in self.times])

    def point_sums(self, k, x_k, kind):
        """(d, N_k): row i = sum over the type-i events strictly before t^k_m of phi_ki / psi_ki."""
        from oracle import functionals as F
        x = np.asarray(x_k, dtype=float)
        imap = OE.interval_map(self.plain, k)
        fn = F.phi if kind == 'phi' else F.psi
        tk = self.times[k]
        out = np.zeros((self.plain['d'], len(tk)))
        for i in range(self.plain['d']):
            ti = self.times[i]
            par = x[imap[i][0]:imap[i][1]]
            for m in range(len(tk)):
                hi = m if i == k else np.searchsorted(ti, tk[m], side='left')
                if hi > 0:
                    out[i][m] = np.sum(fn(self.plain['kernels'][k][i], tk[m] - ti[:hi], par))
        return out

    def exact_max_lag(self, ks, xs, compute_f_sum=True, compute_diff_f_sum=True):
        """Largest window of the oracle's windowed sums (oracle/exact.py phi_window / ups_window)."""
        w = 0.
        d = self.plain['d']
        KU = OE.ups_kernel(self.plain)
        for k, x in zip(ks, xs):
            imap = OE.interval_map(self.plain, k)
            par = [np.asarray(x, dtype=float)[a:b] for a, b in imap]
            for i in range(d):
                w = max(w, OE.phi_window(self.plain['kernels'][k][i], par[i]))
                for j in range(d):
                    w = max(w, OE.ups_window(KU, KU, par[i], par[j]))
        return w

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

    # ---- marked models: per-event weights (aslsd_set_weights / aslsd_exact_sums_weighted) ----
    def set_weights(self, vectors):
        self.weights = [(int(i), np.asarray(v, dtype=float)) for i, v in vectors]

    def exact_sums_weighted(self, k, x_k, roles, term_mask=None, compute_f_sum=True,
                            compute_diff_f_sum=True, ranges=None):
        """Literal O(N^2) weighted sums in the packed layout of exact_sums_packed."""
        from oracle import functionals as F
        self.n_calls += 1
        model, times, T_f = self.plain, self.times, self.T_f
        d = model['d']
        K = model['kernels']
        KU = OE.ups_kernel(model)
        imap = OE.interval_map(model, k)
        roles = np.asarray(roles).reshape(5, d)
        x_k = np.asarray(x_k, dtype=float)

        def wv(role, dim, n):
            v = roles[role][dim]
            if v < 0:
                return np.ones(n)
            assert self.weights[v][0] == dim
            return self.weights[v][1]
        layout, out_len = term_layout(self.desc, k)
        out = np.zeros(out_len)
        mask = np.ones(len(layout), dtype=bool) if term_mask is None else np.asarray(term_mask, dtype=bool)
        for (name, off, n1, n2, n_func), on in zip(layout, mask):
            if not on:
                continue
            kind, rest = name.split('_', 1)
            idx = [int(v) for v in rest.split('_')]
            i = idx[0]
            x_ki = x_k[imap[i][0]:imap[i][1]]
            t2e = T_f - times[i]
            if kind == 'psi':
                w = wv(3, i, len(times[i]))
                lo, hi = (0, len(times[i])) if ranges is None else (int(ranges[i][0]), int(ranges[i][1]))
                s_psi = np.sum(F.psi(K[k][i], t2e[lo:hi], x_ki) * w[lo:hi])
                out[off] = x_k[0] * s_psi
                for p in range(n1):
                    out[off + 1 + p] = x_k[0] * np.sum(F.diff_psi(K[k][i], t2e[lo:hi], p, x_ki) * w[lo:hi])
                out[off + 1 + n1] = s_psi
            elif kind == 'upsilonzero':
                w = wv(4, i, len(times[i]))
                out[off] = np.sum(F.upsilon(KU, KU, t2e, 0., x_ki, x_ki) * w)
                for p in range(n1):
                    out[off + 1 + p] = np.sum(F.diff_sim_upsilon(K[k][i], t2e, 0., p, x_ki) * w)
            elif kind == 'phi':
                tk, ti = times[k], times[i]
                w = wv(0, i, len(ti))
                for m in range(len(tk)):
                    hi = m if i == k else np.searchsorted(ti, tk[m], side='left')
                    if hi == 0:
                        continue
                    s = tk[m] - ti[:hi]
                    out[off] += np.sum(F.phi(K[k][i], s, x_ki) * w[:hi])
                    for p in range(n1):
                        out[off + 1 + p] += np.sum(F.diff_phi(K[k][i], s, p, x_ki) * w[:hi])
            else:
                j = idx[1]
                x_kj = x_k[imap[j][0]:imap[j][1]]
                ti, tj = times[i], times[j]
                wa, wb = wv(1, i, len(ti)), wv(2, j, len(tj))
                for m in range(len(ti)):
                    hi = m if i == j else np.searchsorted(tj, ti[m], side='left')
                    if hi == 0:
                        continue
                    s = ti[m] - tj[:hi]
                    tt = (T_f - ti[m]) * np.ones(hi)
                    loc = wa[m] * wb[:hi]
                    out[off] += np.sum(F.upsilon(KU, KU, tt, s, x_ki, x_kj) * loc)
                    if i == j:
                        for p in range(n1):
                            out[off + 1 + p] += np.sum(F.diff_sim_upsilon(K[k][i], tt, s, p, x_ki) * loc)
                    else:
                        for p in range(n1):
                            out[off + 1 + p] += np.sum(
                                F.diff_cross_upsilon(KU, KU, tt, s, 1, p, x_ki, x_kj) * loc)
                        for p in range(n2):
                            out[off + 1 + n1 + p] += np.sum(
                                F.diff_cross_upsilon(KU, KU, tt, s, 2, p, x_ki, x_kj) * loc)
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

    def exact_packed_len_batch(self, ks):
        return sum(term_layout(self.desc, k)[1] for k in ks)

    def exact_packed_device_batch(self, ks, xs, packed_ptr, compute_f_sum=True, compute_diff_f_sum=True,
                                  ranges=None):
        import ctypes
        n = self.exact_packed_len_batch(ks)
        buf = np.ctypeslib.as_array((ctypes.c_double * n).from_address(packed_ptr))
        pos = 0
        for k, x in zip(ks, xs):
            packed = self.exact_sums_packed(k, x, compute_f_sum, compute_diff_f_sum, ranges=ranges)
            buf[pos:pos + len(packed)] = packed
            pos += len(packed)

    def split_packed(self, ks, packed):
        out, pos = [], 0
        for k in ks:
            n = term_layout(self.desc, k)[1]
            out.append(np.array(packed[pos:pos + n], dtype=float))
            pos += n
        return out

    def exact_assemble_batch(self, ks, xs, raw_host, compute_f_sum=True, compute_diff_f_sum=True):
        stride = len(raw_host) // len(ks)
        return [np.array(raw_host[q * stride:q * stride + term_layout(self.desc, k)[1]], dtype=float)
                for q, k in enumerate(ks)]

    def sampled_sums(self, k, x_k, terms, stratum_offsets, ix_m, ix_n=None, strata_sizes=None,
                     default_alloc=None, compute_diff_f_sum=True, grad_alloc=False, roles=None):
        self.n_calls += 1
        if roles is not None:
            roles = np.asarray(roles).reshape(5, self.plain['d'])

        def weight(kind, i, j, m, n):
            """per-sample weight of a marked model (aslsd_sampled_sums_weighted)"""
            def vec(role, dim, ix):
                v = roles[role][dim]
                return 1. if v < 0 else self.weights[v][1][ix]
            if kind == 0:
                return vec(3, i, m)
            if kind == 1:
                return vec(4, i, m)
            if kind == 2:
                return vec(0, i, n)
            return vec(1, i, m) * vec(2, j, n)
        sums, out_off, means, stds = [], [0], [], []
        s_ix = 0
        groups = []
        for term in terms:
            kind, i, j, n_strata = term[:4]
            sat_t, v_sat = (term[4], term[5]) if len(term) > 4 else (np.inf, 0.)
            group = None
            for s in range(n_strata):
                a, b = stratum_offsets[s_ix], stratum_offsets[s_ix + 1]
                vals = OS.term_values(self.plain, k, x_k, self.times, self.T_f, kind, i, j,
                                      ix_m[a:b], None if ix_n is None else ix_n[a:b])
                if kind < 2:    # saturated samples carry the caller's constant (aslsd_sample_term)
                    far = (self.T_f - self.times[i][ix_m[a:b]]) >= sat_t
                    vals[far, 0] = v_sat
                if roles is not None:
                    w = weight(kind, i, j, ix_m[a:b], None if ix_n is None else ix_n[a:b])
                    vals = vals * (np.asarray(w, dtype=float) * np.ones(b - a))[:, None]
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
