// Batched searchsorted: kappa / varpi / lag_sizes for every ordered pair of
// event types in one launch.  Replaces the per-event Python bisect loops of
// stats/events/time_ordering.py:9-84 (get_kappa_varpi) and :87-119
// (get_lag_sizes).
//
// Closed form:
//   kappa[i][j][n]   = lower_bound(t^i, t^j_n) - 1                 n < N_j
//   kappa[i][j][N_j] = lower_bound(t^i[0:N_i-1], T_f)              (:64-65, no -1)
//   varpi[i][j][h]   = upper_bound(t^i, t^j_{h-1})                 1 <= h <= H_ij
//   H_ij             = kappa[j][i][N_i-1] + 1 = lower_bound(t^j, t^i_{N_i-1})
//   lag_sizes[i][j][h] = N_i - varpi[i][j][h],  [0] = 0
// HBM-bound integer work: every query is read once, every output written once;
// the haystack is reached through the per-type bucket table (common.cuh), whose
// probes of neighbouring (sorted) queries share cache lines.
#include "common.cuh"

struct PairDev {
  int64_t kappa_off;   // offset of kappa[i][j] in the kappa buffer
  int64_t varpi_off;   // offset of varpi[i][j] / lag_sizes[i][j]
  int64_t H;           // number of lags (i != j)
};

__global__ void lag_count_kernel(const DimDev* __restrict__ dims, int d, int64_t* __restrict__ H) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d * d) return;
  int i = p / d, j = p % d;
  if (i == j) { H[p] = dims[i].n - 1; return; }
  const DimDev Di = dims[i], Dj = dims[j];
  H[p] = dim_lower_bound(Dj, Di.t[Di.n - 1]);
}

// IM_Q independent queries per thread, their chains of dependent loads (query -> bucket table ->
// a few probes of the haystack -> tie check) interleaved; the queries of a thread are blockDim
// apart, so every load and store of a warp stays coalesced.  Indices are 32-bit inside the kernel
// (the bucket tables are), the bucket of a query is taken from one multiply without the boundary
// fix-up of dim_lower_bound - the bracket is widened by one bucket on either side instead.
// History (C5, 3.83 GB algorithmic): one chain per thread, 64-bit indices 1.13 ms (52 % of the HBM
// rate, latency-bound); four chains 0.92 ms (63 %, 8.2 warp instructions per query: issue-bound);
// a variant staging the haystack segment of a block of 1024 queries in shared memory 1.06 ms.
#ifndef IM_NT
#define IM_NT 256
#endif
#ifndef IM_Q
#define IM_Q 4
#endif
// the tables are written once and not read back by this kernel: streaming stores keep them from
// displacing the haystacks in L2
#ifdef IM_STREAMING_STORES   // measured slower: 0.86 ms against 0.81 ms at C5
#define IM_ST(ptr, val) __stcs((long long*)(ptr), (long long)(val))
#else
#define IM_ST(ptr, val) (*(ptr) = (val))
#endif

__global__ void __launch_bounds__(IM_NT)
index_maps_kernel(const DimDev* __restrict__ dims, const PairDev* __restrict__ pairs, int d, double T_f,
                  int64_t* __restrict__ kappa, int64_t* __restrict__ varpi,
                  int64_t* __restrict__ lags) {
  const int p = blockIdx.y;
  const int i = p / d, j = p % d;
  const DimDev Di = dims[i], Dj = dims[j];
  const PairDev P = pairs[p];
  int64_t* kap = kappa ? kappa + P.kappa_off : nullptr;
  int64_t* var = varpi ? varpi + P.varpi_off : nullptr;
  int64_t* lag = lags ? lags + P.varpi_off : nullptr;
  if (i == j) {
    // kappa[i][i] = arange(N+1) - 1 (:52-53); varpi[i][i] = arange(N) (:70-71);
    // lag_sizes[i][i][h] = N - h, [0] = 0
    for (int64_t n = (int64_t)blockIdx.x * IM_NT + threadIdx.x; n <= Dj.n; n += (int64_t)gridDim.x * IM_NT) {
      if (kap) IM_ST(&kap[n], n - 1);
      if (n < Di.n) {
        if (var) IM_ST(&var[n], n);
        if (lag) IM_ST(&lag[n], n == 0 ? 0 : Di.n - n);
      }
    }
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // terminal entry: bisect_left(t^i, T_f, 0, N_i - 1) (:64-65, no -1); lag 0
    if (kap) {
      const int64_t q = dim_lower_bound(Di, T_f);
      kap[Dj.n] = q < Di.n - 1 ? q : Di.n - 1;
    }
    if (var) var[0] = 0;
    if (lag) lag[0] = 0;
  }
  const bool want_ub = var || lag;
  const int n_i = (int)Di.n, n_j = (int)Dj.n, n_b = (int)Di.n_buckets, H = (int)min(P.H, (int64_t)n_j);
  const double* __restrict__ ti = Di.t;
  const int32_t* __restrict__ tab = Di.tab;
  const int stride = (int)gridDim.x * IM_NT * IM_Q;
  for (int n0 = (int)blockIdx.x * IM_NT * IM_Q + (int)threadIdx.x; n0 < n_j; n0 += stride) {
    double x[IM_Q];
    int lo[IM_Q], hi[IM_Q];
#pragma unroll
    for (int k = 0; k < IM_Q; ++k) {
      const int n = n0 + k * IM_NT;
      x[k] = n < n_j ? Dj.t[n] : Dj.t[n_j - 1];
    }
#pragma unroll
    for (int k = 0; k < IM_Q; ++k) {
      const double gf = (x[k] - Di.t0) * Di.inv_w;
      const int g = gf <= 0. ? 0 : (gf >= (double)(n_b - 1) ? n_b - 1 : (int)gf);
      // the bucket of x is g - 1, g or g + 1 (rounding of gf): lower bounds of the neighbours as bracket
      lo[k] = g > 0 ? tab[g - 1] : 0;
      hi[k] = tab[min(g + 2, n_b)];
    }
    bool any = true;
    while (any) {
      any = false;
      double v[IM_Q];
      int mid[IM_Q];
#pragma unroll
      for (int k = 0; k < IM_Q; ++k) {
        mid[k] = (lo[k] + hi[k]) >> 1;
        v[k] = ti[min(mid[k], n_i - 1)];
      }
#pragma unroll
      for (int k = 0; k < IM_Q; ++k)
        if (lo[k] < hi[k]) {
          if (v[k] < x[k]) lo[k] = mid[k] + 1; else hi[k] = mid[k];
          any |= lo[k] < hi[k];
        }
    }
    // 'right' side: the element at the lower bound, to step over ties
    double at[IM_Q];
    if (want_ub) {
#pragma unroll
      for (int k = 0; k < IM_Q; ++k) at[k] = ti[min(lo[k], n_i - 1)];
    }
#pragma unroll
    for (int k = 0; k < IM_Q; ++k) {
      const int n = n0 + k * IM_NT;
      if (n >= n_j) continue;
      int q = lo[k];
      if (kap) IM_ST(&kap[n], (int64_t)q - 1);
      if (want_ub && n < H) {
        if (q < n_i && at[k] == x[k]) {
          ++q;
          while (q < n_i && ti[q] == x[k]) ++q;
        }
        if (var) IM_ST(&var[n + 1], (int64_t)q);
        if (lag) IM_ST(&lag[n + 1], (int64_t)(n_i - q));
      }
    }
  }
}

static int compute_lag_counts(aslsd_ctx* ctx, std::vector<int64_t>& H) {
  const int d = ctx->d;
  for (int i = 0; i < d; ++i)
    if (ctx->n_events[i] == 0)
      return aslsd_fail(ctx, ASLSD_ERR_ARG, "index maps: an event type has no events (IndexError upstream)");
  int rc;
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_im_H, &ctx->im_H_cap, sizeof(int64_t) * d * d))) return rc;
  int64_t* dH = ctx->d_im_H;
  lag_count_kernel<<<(d * d + 127) / 128, 128, 0, ctx->stream>>>((const DimDev*)ctx->d_dims, d, dH);
  H.resize(d * d);
  cudaError_t e = cudaMemcpyAsync(H.data(), dH, sizeof(int64_t) * d * d, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  CUDA_TRY(ctx, e);
  return ASLSD_OK;
}

extern "C" int aslsd_index_map_sizes(aslsd_ctx* ctx, int64_t* kappa_len, int64_t* varpi_len) {
  if (!ctx || ctx->d == 0 || !ctx->d_times) return aslsd_fail(ctx, ASLSD_ERR_ARG, "index maps: no events set");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (int vrc = events_verdict(ctx, true)) return vrc;
  std::vector<int64_t> H;
  int rc = compute_lag_counts(ctx, H);
  if (rc) return rc;
  const int d = ctx->d;
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      if (kappa_len) kappa_len[i * d + j] = ctx->n_events[j] + 1;
      if (varpi_len) varpi_len[i * d + j] = (i == j) ? ctx->n_events[i] : H[i * d + j] + 1;
    }
  return ASLSD_OK;
}

extern "C" int aslsd_index_maps(aslsd_ctx* ctx, int64_t* const* kappa_out, int64_t* const* varpi_out,
                                int64_t* const* lag_sizes_out) {
  if (!ctx || ctx->d == 0 || !ctx->d_times) return aslsd_fail(ctx, ASLSD_ERR_ARG, "index maps: no events set");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (int vrc = events_verdict(ctx, true)) return vrc;
  const int d = ctx->d;
  std::vector<int64_t> H;
  int rc = compute_lag_counts(ctx, H);
  if (rc) return rc;
  std::vector<PairDev> pairs(d * d);
  int64_t ktot = 0, vtot = 0, nmax = 0;
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      PairDev& P = pairs[i * d + j];
      P.kappa_off = ktot;
      P.varpi_off = vtot;
      P.H = (i == j) ? 0 : H[i * d + j];
      // every array starts at an even element offset (16-byte stores); the same rule lays out the
      // caller's block when it wants a single read-back per table (Engine.index_maps)
      ktot += (ctx->n_events[j] + 1 + 1) & ~(int64_t)1;
      vtot += (((i == j) ? ctx->n_events[i] : H[i * d + j] + 1) + 1) & ~(int64_t)1;
      nmax = std::max(nmax, ctx->n_events[j] + 1);
    }
  // tables live in buffers of the context (no allocation per call)
  int64_t *dK = nullptr, *dV = nullptr, *dL = nullptr;
  if ((rc = ensure_device_buffer(ctx, &ctx->d_im_pairs, &ctx->im_pairs_cap, sizeof(PairDev) * d * d))) return rc;
  if (kappa_out && (rc = ensure_device_buffer(ctx, (void**)&ctx->d_im_kappa, &ctx->im_kappa_cap, sizeof(int64_t) * ktot)))
    return rc;
  if (varpi_out && (rc = ensure_device_buffer(ctx, (void**)&ctx->d_im_varpi, &ctx->im_varpi_cap, sizeof(int64_t) * vtot)))
    return rc;
  if (lag_sizes_out &&
      (rc = ensure_device_buffer(ctx, (void**)&ctx->d_im_lags, &ctx->im_lags_cap, sizeof(int64_t) * vtot))) return rc;
  PairDev* dP = (PairDev*)ctx->d_im_pairs;
  if (kappa_out) dK = ctx->d_im_kappa;
  if (varpi_out) dV = ctx->d_im_varpi;
  if (lag_sizes_out) dL = ctx->d_im_lags;
  cudaError_t e = cudaMemcpyAsync(dP, pairs.data(), sizeof(PairDev) * d * d, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    const int64_t per_block = (int64_t)IM_NT * IM_Q;
    int64_t bx = (nmax + per_block - 1) / per_block;
    // grid.x: enough blocks to cover the longest query array, capped at a few
    // waves of the 148 SMs (grid-stride loop beyond that)
    dim3 grid((unsigned)std::min<int64_t>(bx, 148 * 32), (unsigned)(d * d));
    cudaEventRecord(ctx->ev0, ctx->stream);
    index_maps_kernel<<<grid, IM_NT, 0, ctx->stream>>>((const DimDev*)ctx->d_dims, dP, d, ctx->T_f, dK, dV, dL);
    cudaEventRecord(ctx->ev1, ctx->stream);
    e = cudaGetLastError();
  }
  // read-back: one transfer per table when the caller's d x d arrays lie back to back in one block
  // (Engine.index_maps allocates them that way), one per array otherwise
  auto contiguous = [&](int64_t* const* out, bool is_kappa) {
    if (!out) return false;
    for (int p = 0; p < d * d; ++p) {
      if (!out[p]) return false;
      const int64_t off = is_kappa ? pairs[p].kappa_off : pairs[p].varpi_off;
      if (out[p] != out[0] + off) return false;
    }
    return true;
  };
  struct Table { int64_t* const* out; int64_t* dev; bool is_kappa; int64_t total; };
  const Table tables[3] = {{kappa_out, dK, true, ktot}, {varpi_out, dV, false, vtot}, {lag_sizes_out, dL, false, vtot}};
  for (const Table& T : tables) {
    if (!T.out || e != cudaSuccess) continue;
    if (contiguous(T.out, T.is_kappa)) {
      e = cudaMemcpyAsync(T.out[0], T.dev, sizeof(int64_t) * T.total, cudaMemcpyDeviceToHost, ctx->stream);
      continue;
    }
    for (int p = 0; p < d * d && e == cudaSuccess; ++p) {
      if (!T.out[p]) continue;
      const int i = p / d, j = p % d;
      const int64_t len = T.is_kappa ? ctx->n_events[j] + 1 : ((i == j) ? ctx->n_events[i] : H[p] + 1);
      const int64_t off = T.is_kappa ? pairs[p].kappa_off : pairs[p].varpi_off;
      e = cudaMemcpyAsync(T.out[p], T.dev + off, sizeof(int64_t) * len, cudaMemcpyDeviceToHost, ctx->stream);
    }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms = ms;
    ctx->last_launches = 1;
  }
  CUDA_TRY(ctx, e);
  return ASLSD_OK;
}
