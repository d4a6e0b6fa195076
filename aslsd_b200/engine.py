"""Engine: one device context of the B200 library, behind plain numpy arguments.

This is the thin layer between the reference-facing host classes
(``aslsd_b200.events``, ``aslsd_b200.estimators``) and the C ABI.  One Engine
may be shared by the d estimators of a fit (same ProcessPath).
"""
import ctypes

import numpy as np

from aslsd_b200 import capi
from aslsd_b200.model import describe_model


class EngineError(RuntimeError):
    pass


def term_layout(desc, k):
    """Packed output layout of aslsd_exact_sums: list of
    (log_name, offset, n1, n2, n_func) in the order the reference evaluates the
    terms (mhp_exact_estimator.py:722-756); n_func as in make_param_indexing
    (:141-170)."""
    cache = getattr(desc, '_term_layouts', None)      # (an estimator call asks twice; d = 10: 120 terms)
    if cache is None:
        try:
            cache = desc._term_layouts = {}
        except AttributeError:
            cache = {}
    if k in cache:
        return cache[k]
    d = desc.d
    npar = desc.matrix_n_param[k]
    out, pos = [], 0
    for i in range(d):
        for name, n1, n2, n_func in (('psi_%d' % i, npar[i], 1, 2),
                                     ('upsilonzero_%d' % i, npar[i], 0, 1),
                                     ('phi_%d' % i, npar[i], 0, 1),
                                     ('upsilon_%d_%d' % (i, i), npar[i], 0, 1)):
            out.append((name, pos, n1, n2, n_func))
            pos += 1 + n1 + n2
    for i in range(d):
        for j in range(d):
            if i != j:
                out.append(('upsilon_%d_%d' % (i, j), pos, npar[i], npar[j], 2))
                pos += 1 + npar[i] + npar[j]
    cache[k] = (out, pos)
    return out, pos


def unpack_terms(desc, k, packed):
    """Packed vector -> (sum_f, sum_diff_f) dictionaries with the shapes of
    make_null_sumarray (mhp_exact_estimator.py:242-255)."""
    layout, _ = term_layout(desc, k)
    sum_f, sum_diff_f = {}, {}
    for name, off, n1, n2, n_func in layout:
        sum_f[name] = float(packed[off])
        if n_func == 2:
            sum_diff_f[name] = [None, packed[off + 1:off + 1 + n1].copy(),
                                packed[off + 1 + n1:off + 1 + n1 + n2].copy()]
        else:
            sum_diff_f[name] = packed[off + 1:off + 1 + n1].copy()
    return sum_f, sum_diff_f


class Engine:
    def __init__(self, device=0):
        self._lib = capi.load_library()
        self._ctx = ctypes.c_void_p()
        rc = self._lib.aslsd_create(ctypes.byref(self._ctx), int(device))
        if rc != 0 or not self._ctx:
            raise EngineError('aslsd_create failed (rc=%d): no usable CUDA device %d; '
                              'the B200 engine has no CPU fallback' % (rc, device))
        self.device = int(device)
        self.desc = None
        self.d = None
        self.n_events = None
        self.T_f = None

    def close(self):
        if getattr(self, '_ctx', None):
            self._lib.aslsd_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.aslsd_last_error(self._ctx)
            msg = msg.decode() if msg else ''
            if rc == capi.ERR_UNSUPPORTED:
                raise NotImplementedError(msg)
            raise EngineError('aslsd_b200 error %d: %s' % (rc, msg))

    def set_stream(self, cuda_stream):
        """Run all subsequent work on the caller's CUDA stream (a raw cudaStream_t
        handle, e.g. ``torch.cuda.current_stream().cuda_stream``); ``None`` restores
        the context's own stream.  Handle 0 is torch's default stream: it is mapped
        to cudaStreamLegacy (0x1), since NULL means "own stream" in the C ABI."""
        if cuda_stream is None:
            handle = 0
        else:
            handle = int(cuda_stream) or 1
        self._check(self._lib.aslsd_set_stream(self._ctx, ctypes.c_void_p(handle)))

    def synchronize(self):
        """Wait for everything enqueued on the context's stream."""
        self._check(self._lib.aslsd_synchronize(self._ctx))

    def last_timing(self):
        ms = ctypes.c_double()
        n = ctypes.c_int64()
        self._check(self._lib.aslsd_last_timing(self._ctx, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def last_moment_timing(self):
        """Duration (ms) of the lag-moment kernel of the last exact-sum call."""
        ms = ctypes.c_double()
        self._check(self._lib.aslsd_last_moment_timing(self._ctx, ctypes.byref(ms)))
        return ms.value

    def set_option(self, option, value):
        self._check(self._lib.aslsd_set_option(self._ctx, int(option), int(value)))

    def use_lag_quadrature(self, enabled=True):
        """Lag quadrature for the interior targets of every double sum (default on);
        off: every row goes through the per-item pair kernels."""
        self.set_option(capi.OPT_LAG_QUADRATURE, 1 if enabled else 0)

    def use_weight_cache(self, enabled=True):
        """Keep the quadrature weights between calls on the same plan and events (default
        on): the iterations of a fit with fixed (beta, delta) then skip the pass over the
        event pairs.  Off: every call recomputes them (benchmark setting)."""
        self.set_option(capi.OPT_WEIGHT_CACHE, 1 if enabled else 0)

    def use_async_events(self, enabled=True):
        """Page-locked event arrays are uploaded in place and checked for order on the device.
        Off (default): ``set_events`` waits for that verdict and raises on unsorted times, as it
        does for pageable arrays.  On: it returns once the transfer is enqueued, so the planning
        of the next call overlaps the transfer; unsorted times then raise at the next call that
        waits for the device."""
        self.set_option(capi.OPT_ASYNC_EVENTS, 1 if enabled else 0)

    def last_work_counts(self):
        """Work of the last exact-sum call (see aslsd_last_work_counts)."""
        c = np.zeros(8, dtype=np.int64)
        self._check(self._lib.aslsd_last_work_counts(self._ctx, capi.ptr_i64(c), 8))
        names = ('pair_evals', 'single_evals', 'moment_pairs', 'quad_evals', 'streams', 'bins',
                 'rows', 'items')
        return dict(zip(names, (int(v) for v in c)))

    def last_eval_counts(self):
        """(pair-kernel, single-kernel) closed-form evaluations of the last
        exact-sum call."""
        a, b = ctypes.c_int64(), ctypes.c_int64()
        self._check(self._lib.aslsd_last_eval_counts(self._ctx, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def measure_fp64_peak(self):
        """Measured DFMA throughput in TFLOP/s (roofline denominator)."""
        v = ctypes.c_double()
        self._check(self._lib.aslsd_measure_fp64_peak(self._ctx, ctypes.byref(v)))
        return v.value

    # ---- events ----------------------------------------------------------------------
    def set_events(self, list_times, T_f):
        times = [capi.as_f64(t) for t in list_times]
        d = len(times)
        n_events = np.array([len(t) for t in times], dtype=np.int64)
        ptrs = (capi.c_f64_p * d)(*[capi.ptr_f64(t) for t in times])
        self._check(self._lib.aslsd_set_events(self._ctx, d, capi.ptr_i64(n_events), ptrs,
                                               float(T_f)))
        self.d, self.n_events, self.T_f = d, n_events, float(T_f)
        self._events_key = None
        # page-locked arrays are read in place by the library (no staging copy): keep them alive
        self._events_arrays = times

    def set_horizon(self, T_f):
        """Move T_f of the events on the device (aslsd_set_horizon): later sums see T_f - t_m at the
        new horizon; the caller restricts the targets to the events before it (``ranges``)."""
        self._check(self._lib.aslsd_set_horizon(self._ctx, float(T_f)))
        self.T_f = float(T_f)

    def index_maps(self, want_kappa=True, want_varpi=True, want_lag_sizes=True):
        """kappa, varpi, lag_sizes as d x d lists of int64 arrays
        (time_ordering.py:9-119)."""
        d = self.d
        klen = np.zeros(d * d, dtype=np.int64)
        vlen = np.zeros(d * d, dtype=np.int64)
        self._check(self._lib.aslsd_index_map_sizes(self._ctx, capi.ptr_i64(klen),
                                                    capi.ptr_i64(vlen)))

        def alloc(lens, want):
            # the d x d arrays of a table are views of one block, back to back in pair order:
            # the library then reads the whole table back with a single transfer
            if not want:
                return None, None
            # (every array at an even element offset, the layout of the device tables)
            padded = (np.asarray(lens, dtype=np.int64) + 1) & ~np.int64(1)
            block = np.empty(int(np.sum(padded)), dtype=np.int64)
            cuts = np.concatenate([[0], np.cumsum(padded)]).astype(np.int64)
            arrs = [block[cuts[p]:cuts[p] + int(lens[p])] for p in range(d * d)]
            ptrs = (capi.c_i64_p * (d * d))(*[capi.ptr_i64(a) for a in arrs])
            return arrs, ptrs
        kap, kap_p = alloc(klen, want_kappa)
        var, var_p = alloc(vlen, want_varpi)
        lag, lag_p = alloc(vlen, want_lag_sizes)
        self._check(self._lib.aslsd_index_maps(self._ctx, kap_p, var_p, lag_p))

        def nest(arrs):
            return None if arrs is None else [[arrs[i * d + j] for j in range(d)]
                                              for i in range(d)]
        return nest(kap), nest(var), nest(lag)

    def index_maps_kernel_time(self):
        """(kernel ms, algorithmic bytes) of one index-map launch with the three tables left on
        the device (roofline accounting of the HBM-bound searchsorted kernel: per ordered pair
        i != j read 8 (N_i + N_j), write 8 (N_j + 1) + 16 (H + 1) bytes)."""
        d = self.d
        klen = np.zeros(d * d, dtype=np.int64)
        vlen = np.zeros(d * d, dtype=np.int64)
        self._check(self._lib.aslsd_index_map_sizes(self._ctx, capi.ptr_i64(klen), capi.ptr_i64(vlen)))
        nulls = (capi.c_i64_p * (d * d))()
        self._check(self._lib.aslsd_index_maps(self._ctx, nulls, nulls, nulls))
        ms, _ = self.last_timing()
        n = self.n_events
        nbytes = 0
        for i in range(d):
            for j in range(d):
                nbytes += 8 * int(klen[i * d + j]) + 16 * int(vlen[i * d + j])
                if i != j:
                    nbytes += 8 * int(n[i] + n[j])
        return ms, nbytes

    # ---- simulation --------------------------------------------------------------------
    def simulate(self, model, mu, kernel_param, T_f, seed=1234):
        """Cluster simulation of the MHP ``model`` on the device (aslsd_simulate; the generation
        loop of MHP.simulate, mhp.py:733-854): ``kernel_param[i][j]`` are the free parameters of
        phi_ij as in the reference, ``mu`` the baselines.  Returns ``list_times``.  Same law as
        the reference's simulator, not its random numbers."""
        desc = describe_model(model)
        d = desc.d
        flat = []
        n_basis = (ctypes.c_int32 * (d * d))()
        for i in range(d):
            for j in range(d):
                params = iter(np.asarray(kernel_param[i][j], dtype=np.float64))
                n_basis[i * d + j] = len(desc.bases[i][j])
                for type_name, free_mask, fixed in desc.bases[i][j]:
                    nv = 2 if type_name == 'ExponentialKernel' else 3
                    flat.append((type_name, [float(next(params)) if (free_mask >> q) & 1 else float(fixed[q])
                                             for q in range(nv)]))
        arr = (capi.Basis * max(1, len(flat)))()
        for ix, (type_name, v) in enumerate(flat):
            arr[ix].type = capi.BASIS_EXPONENTIAL if type_name == 'ExponentialKernel' else capi.BASIS_GAUSSIAN
            arr[ix].free_mask = 0
            for q, val in enumerate(v):
                arr[ix].fixed[q] = val
        mu = capi.as_f64(mu)
        counts = np.zeros(d, dtype=np.int64)
        self._check(self._lib.aslsd_simulate(self._ctx, d, n_basis, arr, capi.ptr_f64(mu), float(T_f),
                                             ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), capi.ptr_i64(counts)))
        times = [np.empty(int(n), dtype=np.float64) for n in counts]
        ptrs = (capi.c_f64_p * d)(*[capi.ptr_f64(t) for t in times])
        self._check(self._lib.aslsd_simulate_fetch(self._ctx, ptrs))
        return times

    # ---- model -----------------------------------------------------------------------
    def set_model(self, model):
        desc = describe_model(model)
        n_basis, bases = desc.c_arrays()
        self._check(self._lib.aslsd_set_model(self._ctx, desc.d, n_basis, bases))
        self.desc = desc
        self._model_ref = None
        return desc

    # ---- exact sums ------------------------------------------------------------------
    def _ranges(self, ranges):
        if ranges is None:
            return None, None, None, None
        lo = np.array([r[0] for r in ranges], dtype=np.int64)
        hi = np.array([r[1] for r in ranges], dtype=np.int64)
        return lo, hi, capi.ptr_i64(lo), capi.ptr_i64(hi)

    def exact_sums_packed(self, k, x_k, compute_f_sum=True, compute_diff_f_sum=True,
                          ranges=None, count_pairs=False):
        x = capi.as_f64(np.asarray(x_k, dtype=np.float64))
        if len(x) != self.desc.n_param_k(k):
            raise ValueError('x_k has %d entries, expected %d' % (len(x), self.desc.n_param_k(k)))
        layout, out_len = term_layout(self.desc, k)
        out = np.zeros(out_len, dtype=np.float64)
        flags = (capi.ASLSD_F if compute_f_sum else 0) | (capi.ASLSD_DIFF if compute_diff_f_sum else 0)
        lo, hi, lo_p, hi_p = self._ranges(ranges)
        counts = np.zeros(len(layout), dtype=np.int64) if count_pairs else None
        self._check(self._lib.aslsd_exact_sums(
            self._ctx, int(k), capi.ptr_f64(x), flags, lo_p, hi_p, capi.ptr_f64(out),
            capi.ptr_i64(counts) if count_pairs else None))
        if count_pairs:
            return out, {entry[0]: int(c) for entry, c in zip(layout, counts)}
        return out

    def exact_sums(self, k, x_k, compute_f_sum=True, compute_diff_f_sum=True, ranges=None,
                   count_pairs=False):
        res = self.exact_sums_packed(k, x_k, compute_f_sum, compute_diff_f_sum, ranges,
                                     count_pairs)
        packed, counts = res if count_pairs else (res, None)
        sum_f, sum_diff_f = unpack_terms(self.desc, k, packed)
        if count_pairs:
            return sum_f, sum_diff_f, counts
        return sum_f, sum_diff_f

    # ---- per-event sums (intensity at the jumps, compensator) --------------------------
    def point_sums(self, k, x_k, kind):
        """(d, n_events[k]) array: row i = sum over the earlier type-i events of phi_ki
        (kind 'phi') or psi_ki (kind 'psi') at every event of type k."""
        x = capi.as_f64(np.asarray(x_k, dtype=np.float64))
        if len(x) != self.desc.n_param_k(k):
            raise ValueError('x_k has %d entries, expected %d' % (len(x), self.desc.n_param_k(k)))
        out = np.zeros((self.d, int(self.n_events[k])), dtype=np.float64)
        self._check(self._lib.aslsd_point_sums(self._ctx, int(k), capi.ptr_f64(x),
                                               {'phi': 0, 'psi': 1}[kind], capi.ptr_f64(out)))
        return out

    # ---- marked models: per-event weights ---------------------------------------------
    def set_weights(self, vectors):
        """Upload per-event weight vectors: ``vectors`` is a list of (event type, values)
        with one value per event of that type.  They are referred to by position in
        ``exact_sums_weighted``."""
        vals = [capi.as_f64(np.asarray(v, dtype=np.float64)) for _, v in vectors]
        dims = np.ascontiguousarray([int(i) for i, _ in vectors], dtype=np.int32)
        for (i, _), v in zip(vectors, vals):
            if len(v) != self.n_events[i]:
                raise ValueError('weight vector of type %d has %d entries, expected %d'
                                 % (i, len(v), self.n_events[i]))
        ptrs = (capi.c_f64_p * max(1, len(vals)))(*[capi.ptr_f64(v) for v in vals])
        self._check(self._lib.aslsd_set_weights(self._ctx, len(vals),
                                                dims.ctypes.data_as(capi.c_i32_p), ptrs))

    def exact_sums_weighted(self, k, x_k, roles, term_mask=None, compute_f_sum=True,
                            compute_diff_f_sum=True, ranges=None):
        """Packed sums of ``exact_sums_packed`` with per-event weights
        (aslsd_exact_sums_weighted): ``roles`` is a (5, d) integer array of weight
        vector positions (-1: unweighted); ``term_mask`` an optional boolean array over
        the terms in output order."""
        x = capi.as_f64(np.asarray(x_k, dtype=np.float64))
        if len(x) != self.desc.n_param_k(k):
            raise ValueError('x_k has %d entries, expected %d' % (len(x), self.desc.n_param_k(k)))
        layout, out_len = term_layout(self.desc, k)
        roles = np.ascontiguousarray(roles, dtype=np.int32).reshape(5 * self.d)
        mask = mask_p = None
        if term_mask is not None:
            mask = np.ascontiguousarray(term_mask, dtype=np.uint8)
            if len(mask) != len(layout):
                raise ValueError('term_mask must have one entry per term')
            mask_p = mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        out = np.zeros(out_len, dtype=np.float64)
        flags = (capi.ASLSD_F if compute_f_sum else 0) | (capi.ASLSD_DIFF if compute_diff_f_sum else 0)
        lo, hi, lo_p, hi_p = self._ranges(ranges)
        self._check(self._lib.aslsd_exact_sums_weighted(
            self._ctx, int(k), capi.ptr_f64(x), flags, lo_p, hi_p,
            roles.ctypes.data_as(capi.c_i32_p), mask_p, capi.ptr_f64(out), None))
        return out

    # ---- sampled sums ----------------------------------------------------------------
    def sampled_n_out(self, k, kind, i, j):
        n = ctypes.c_int32()
        self._check(self._lib.aslsd_sampled_n_out(self._ctx, int(k), int(kind), int(i), int(j),
                                                  ctypes.byref(n)))
        return n.value

    def sampled_sums(self, k, x_k, terms, stratum_offsets, ix_m, ix_n=None, strata_sizes=None,
                     default_alloc=None, compute_diff_f_sum=True, grad_alloc=False, roles=None):
        """One launch for every stratum of every term of a compute_objective call
        (mhp_stratified_estimator.py:1333-1386).

        terms: list of (kind, i, j, n_strata[, sat_t, v_sat]); the strata of all terms are
        concatenated and stratum s owns samples [stratum_offsets[s],
        stratum_offsets[s+1]) of ix_m / ix_n.  Returns (sums, out_offsets,
        stat_mean, stat_std, rel_alloc): sums[out_offsets[s]:out_offsets[s+1]] are
        the plain sums of f and of its derivatives over stratum s."""
        x = capi.as_f64(np.asarray(x_k, dtype=np.float64))
        if len(x) != self.desc.n_param_k(k):
            raise ValueError('x_k has %d entries, expected %d' % (len(x), self.desc.n_param_k(k)))
        n_terms = len(terms)
        arr = (capi.SampleTerm * max(n_terms, 1))()
        n_strata, out_offsets = 0, [0]
        for q, term in enumerate(terms):
            kind, i, j, ns = term[:4]
            arr[q].kind, arr[q].i, arr[q].j, arr[q].n_strata = int(kind), int(i), int(j), int(ns)
            # optional (sat_t, v_sat): the reference's constant value of saturated single-index samples
            arr[q].sat_t, arr[q].v_sat = (float(term[4]), float(term[5])) if len(term) > 4 else (np.inf, 0.)
            n_out = self.sampled_n_out(k, kind, i, j)
            for _ in range(int(ns)):
                out_offsets.append(out_offsets[-1] + n_out)
            n_strata += int(ns)
        offs = np.ascontiguousarray(stratum_offsets, dtype=np.int64)
        if len(offs) != n_strata + 1:
            raise ValueError('stratum_offsets must have n_strata + 1 entries')
        m = np.ascontiguousarray(ix_m, dtype=np.int64)
        n = None if ix_n is None else np.ascontiguousarray(ix_n, dtype=np.int64)
        if len(m) != offs[-1] or (n is not None and len(n) != len(m)):
            raise ValueError('sample arrays do not match stratum_offsets')
        sums = np.zeros(out_offsets[-1], dtype=np.float64)
        stat_mean = np.zeros(n_strata, dtype=np.float64)
        stat_std = np.zeros(n_strata, dtype=np.float64)
        with_alloc = strata_sizes is not None
        sizes = capi.as_f64(strata_sizes) if with_alloc else None
        dflt = capi.as_f64(default_alloc) if with_alloc else None
        rel = np.zeros(n_strata, dtype=np.float64) if with_alloc else None
        flags = capi.ASLSD_F | (capi.ASLSD_DIFF if compute_diff_f_sum else 0) | \
            (capi.ASLSD_GRAD_ALLOC if grad_alloc else 0)
        if roles is None:
            self._check(self._lib.aslsd_sampled_sums(
                self._ctx, int(k), capi.ptr_f64(x), flags, n_terms, arr, capi.ptr_i64(offs),
                capi.ptr_i64(m), capi.ptr_i64(n) if n is not None else None,
                capi.ptr_f64(sizes) if with_alloc else None,
                capi.ptr_f64(dflt) if with_alloc else None,
                capi.ptr_f64(sums), capi.ptr_f64(stat_mean), capi.ptr_f64(stat_std),
                capi.ptr_f64(rel) if with_alloc else None))
        else:
            # marked models: per-event weights by role (see exact_sums_weighted)
            roles = np.ascontiguousarray(roles, dtype=np.int32).reshape(5 * self.d)
            self._check(self._lib.aslsd_sampled_sums_weighted(
                self._ctx, int(k), capi.ptr_f64(x), flags, n_terms, arr, capi.ptr_i64(offs),
                capi.ptr_i64(m), capi.ptr_i64(n) if n is not None else None,
                capi.ptr_f64(sizes) if with_alloc else None,
                capi.ptr_f64(dflt) if with_alloc else None, roles.ctypes.data_as(capi.c_i32_p),
                capi.ptr_f64(sums), capi.ptr_f64(stat_mean), capi.ptr_f64(stat_std),
                capi.ptr_f64(rel) if with_alloc else None))
        return sums, np.asarray(out_offsets, dtype=np.int64), stat_mean, stat_std, rel

    def plan_rows(self):
        """(meta (n_rows, 7) int32: family, target type, source type, same-type flag, component mask, weight
        vectors of targets / sources;
        params (n_rows, 11): beta_1, delta_1, beta_2, delta_2, beta_own, delta_own, s_lo, s_hi, t_int, m_lo,
        m_hi) of the rows of the
        raw accumulator array of the last batch plan (aslsd_exact_plan_rows; host-side diagnostic)."""
        n = ctypes.c_int64(0)
        self._check(self._lib.aslsd_exact_plan_rows(self._ctx, 0, ctypes.byref(n), None, None))
        meta = np.zeros((n.value, 7), dtype=np.int32)
        params = np.zeros((n.value, 11), dtype=np.float64)
        self._check(self._lib.aslsd_exact_plan_rows(self._ctx, n.value, ctypes.byref(n),
                                                    meta.ctypes.data_as(capi.c_i32_p), capi.ptr_f64(params)))
        return meta, params

    def update_allocation(self, strata_sizes, stat_std, default_alloc):
        sizes, std, dflt = (capi.as_f64(a) for a in (strata_sizes, stat_std, default_alloc))
        out = np.zeros(len(sizes), dtype=np.float64)
        self._check(self._lib.aslsd_update_allocation(self._ctx, len(sizes), capi.ptr_f64(sizes),
                                                      capi.ptr_f64(std), capi.ptr_f64(dflt),
                                                      capi.ptr_f64(out)))
        return out

    # ---- binding to a path / model (shared by the d estimators of a fit) -------------
    def bind(self, process_path, model):
        """Upload the events of ``process_path`` and describe ``model`` unless this
        engine already holds them (the d estimators of a fit share one engine)."""
        key = (id(process_path.list_times), float(process_path.T_f),
               tuple(len(t) for t in process_path.list_times))
        if getattr(self, '_events_key', None) != key:
            self.set_events(process_path.list_times, process_path.T_f)
            self._events_key = key
            self._events_ref = process_path.list_times   # keeps id() unique
        if getattr(self, '_model_ref', None) is not model:
            self.set_model(model)
            self._model_ref = model
        return self.desc

    def _batch_args(self, ks, xs):
        xs = [capi.as_f64(np.asarray(x, dtype=np.float64)) for x in xs]
        for k, x in zip(ks, xs):
            if len(x) != self.desc.n_param_k(k):
                raise ValueError('x_k has %d entries, expected %d' % (len(x), self.desc.n_param_k(k)))
        k_arr = np.ascontiguousarray(ks, dtype=np.int32)
        x_ptrs = (capi.c_f64_p * len(xs))(*[capi.ptr_f64(x) for x in xs])
        return xs, k_arr, k_arr.ctypes.data_as(capi.c_i32_p), x_ptrs

    def exact_sums_batch(self, ks, xs, compute_f_sum=True, compute_diff_f_sum=True, ranges=None,
                         count_pairs=False):
        """Packed sums of several sub-problems in one call; work items shared by
        several of them (fixed beta / delta) are evaluated once."""
        xs, k_arr, k_ptr, x_ptrs = self._batch_args(ks, xs)
        outs = [np.zeros(term_layout(self.desc, k)[1], dtype=np.float64) for k in ks]
        out_ptrs = (capi.c_f64_p * len(ks))(*[capi.ptr_f64(o) for o in outs])
        flags = (capi.ASLSD_F if compute_f_sum else 0) | (capi.ASLSD_DIFF if compute_diff_f_sum else 0)
        lo, hi, lo_p, hi_p = self._ranges(ranges)
        cnts = cnt_ptrs = None
        if count_pairs:
            cnts = [np.zeros(len(term_layout(self.desc, k)[0]), dtype=np.int64) for k in ks]
            cnt_ptrs = (capi.c_i64_p * len(ks))(*[capi.ptr_i64(c) for c in cnts])
        self._check(self._lib.aslsd_exact_sums_batch(self._ctx, len(ks), k_ptr, x_ptrs, flags, lo_p,
                                                     hi_p, out_ptrs, cnt_ptrs))
        if count_pairs:
            return outs, [{e[0]: int(c) for e, c in zip(term_layout(self.desc, k)[0], cnt)}
                          for k, cnt in zip(ks, cnts)]
        return outs

    def exact_max_lag(self, ks, xs, compute_f_sum=True, compute_diff_f_sum=True):
        """Largest lag any double sum of these sub-problems reaches (halo of a time block)."""
        xs, k_arr, k_ptr, x_ptrs = self._batch_args(ks, xs)
        flags = (capi.ASLSD_F if compute_f_sum else 0) | (capi.ASLSD_DIFF if compute_diff_f_sum else 0)
        w = ctypes.c_double()
        self._check(self._lib.aslsd_exact_max_lag(self._ctx, len(ks), k_ptr, x_ptrs, flags, ctypes.byref(w)))
        return w.value

    def exact_raw_used(self):
        """Doubles of the raw buffer written by the last raw_device call."""
        n = ctypes.c_int64()
        self._check(self._lib.aslsd_exact_raw_used(self._ctx, ctypes.byref(n)))
        return n.value

    def exact_raw_len_batch(self, n_k):
        n = ctypes.c_int64()
        self._check(self._lib.aslsd_exact_raw_len_batch(self._ctx, int(n_k), ctypes.byref(n)))
        return n.value

    def exact_raw_device_batch(self, ks, xs, raw_device_ptr, compute_f_sum=True,
                               compute_diff_f_sum=True, ranges=None):
        xs, k_arr, k_ptr, x_ptrs = self._batch_args(ks, xs)
        flags = (capi.ASLSD_F if compute_f_sum else 0) | (capi.ASLSD_DIFF if compute_diff_f_sum else 0)
        lo, hi, lo_p, hi_p = self._ranges(ranges)
        self._check(self._lib.aslsd_exact_raw_device_batch(
            self._ctx, len(ks), k_ptr, x_ptrs, flags, lo_p, hi_p, ctypes.c_void_p(raw_device_ptr)))

    def exact_packed_len_batch(self, ks):
        k_arr = np.ascontiguousarray(ks, dtype=np.int32)
        n = ctypes.c_int64()
        self._check(self._lib.aslsd_exact_packed_len_batch(self._ctx, len(ks), k_arr.ctypes.data_as(capi.c_i32_p),
                                                           ctypes.byref(n)))
        return n.value

    def exact_packed_device_batch(self, ks, xs, packed_device_ptr, compute_f_sum=True,
                                  compute_diff_f_sum=True, ranges=None):
        """Kernels + device-side assembly: the packed sums of the sub-problems ``ks`` (concatenated)
        are left at ``packed_device_ptr`` on the engine's stream (asynchronous)."""
        xs, k_arr, k_ptr, x_ptrs = self._batch_args(ks, xs)
        flags = (capi.ASLSD_F if compute_f_sum else 0) | (capi.ASLSD_DIFF if compute_diff_f_sum else 0)
        lo, hi, lo_p, hi_p = self._ranges(ranges)
        self._check(self._lib.aslsd_exact_packed_device_batch(
            self._ctx, len(ks), k_ptr, x_ptrs, flags, lo_p, hi_p, ctypes.c_void_p(packed_device_ptr)))

    def split_packed(self, ks, packed):
        """Concatenated packed sums -> one array per sub-problem."""
        lens = self._packed_lens(ks)
        out, pos = [], 0
        for n in lens:
            out.append(np.array(packed[pos:pos + n], dtype=np.float64))
            pos += n
        return out

    def _packed_lens(self, ks):
        """Lengths of the packed sums of the sub-problems ``ks`` (kept per model description)."""
        cache = getattr(self, '_packed_len_cache', None)
        if cache is None or cache[0] is not self.desc:
            cache = (self.desc, {})
            self._packed_len_cache = cache
        for k in ks:
            if k not in cache[1]:
                cache[1][k] = term_layout(self.desc, k)[1]
        return [cache[1][k] for k in ks]

    def use_device_assembly(self, enabled=True):
        self.set_option(capi.OPT_DEVICE_ASSEMBLY, 1 if enabled else 0)

    def exact_assemble_batch(self, ks, xs, raw_host, compute_f_sum=True, compute_diff_f_sum=True):
        xs, k_arr, k_ptr, x_ptrs = self._batch_args(ks, xs)
        raw = capi.as_f64(raw_host)
        outs = [np.zeros(term_layout(self.desc, k)[1], dtype=np.float64) for k in ks]
        out_ptrs = (capi.c_f64_p * len(ks))(*[capi.ptr_f64(o) for o in outs])
        flags = (capi.ASLSD_F if compute_f_sum else 0) | (capi.ASLSD_DIFF if compute_diff_f_sum else 0)
        self._check(self._lib.aslsd_exact_assemble_batch(self._ctx, len(ks), k_ptr, x_ptrs, flags,
                                                         capi.ptr_f64(raw), out_ptrs))
        return outs

    def exact_assemble_map_batch(self, ks, xs, raw_host, compute_f_sum=True, compute_diff_f_sum=True):
        """Diagnostic: the device assembly's map applied on the host (aslsd_exact_assemble_map_batch)."""
        xs, k_arr, k_ptr, x_ptrs = self._batch_args(ks, xs)
        raw = capi.as_f64(raw_host)
        outs = [np.zeros(term_layout(self.desc, k)[1], dtype=np.float64) for k in ks]
        out_ptrs = (capi.c_f64_p * len(ks))(*[capi.ptr_f64(o) for o in outs])
        flags = (capi.ASLSD_F if compute_f_sum else 0) | (capi.ASLSD_DIFF if compute_diff_f_sum else 0)
        self._check(self._lib.aslsd_exact_assemble_map_batch(self._ctx, len(ks), k_ptr, x_ptrs, flags,
                                                             capi.ptr_f64(raw), out_ptrs))
        return outs

    # split form (multi-GPU): raw accumulators stay on the device
    def exact_raw_len(self, k):
        n = ctypes.c_int64()
        self._check(self._lib.aslsd_exact_raw_len(self._ctx, int(k), ctypes.byref(n)))
        return n.value

    def exact_raw_device(self, k, x_k, raw_device_ptr, compute_f_sum=True,
                         compute_diff_f_sum=True, ranges=None):
        x = capi.as_f64(np.asarray(x_k, dtype=np.float64))
        flags = (capi.ASLSD_F if compute_f_sum else 0) | (capi.ASLSD_DIFF if compute_diff_f_sum else 0)
        lo, hi, lo_p, hi_p = self._ranges(ranges)
        self._check(self._lib.aslsd_exact_raw_device(
            self._ctx, int(k), capi.ptr_f64(x), flags, lo_p, hi_p,
            ctypes.c_void_p(raw_device_ptr), None))

    def exact_assemble(self, k, x_k, raw_host, compute_f_sum=True, compute_diff_f_sum=True):
        x = capi.as_f64(np.asarray(x_k, dtype=np.float64))
        raw = capi.as_f64(raw_host)
        _, out_len = term_layout(self.desc, k)
        out = np.zeros(out_len, dtype=np.float64)
        flags = (capi.ASLSD_F if compute_f_sum else 0) | (capi.ASLSD_DIFF if compute_diff_f_sum else 0)
        self._check(self._lib.aslsd_exact_assemble(self._ctx, int(k), capi.ptr_f64(x), flags,
                                                   capi.ptr_f64(raw), capi.ptr_f64(out)))
        return out
