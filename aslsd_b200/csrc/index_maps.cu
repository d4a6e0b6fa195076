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

__global__ void __launch_bounds__(256)
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
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n <= Dj.n; n += stride) {
    if (i == j) {
      // kappa[i][i] = arange(N+1) - 1 (:52-53); varpi[i][i] = arange(N) (:70-71);
      // lag_sizes[i][i][h] = N - h, [0] = 0
      if (kap) kap[n] = n - 1;
      if (n < Di.n) {
        if (var) var[n] = n;
        if (lag) lag[n] = n == 0 ? 0 : Di.n - n;
      }
      continue;
    }
    if (n == Dj.n) {
      // terminal entry: bisect_left(t^i, T_f, 0, N_i - 1)
      if (kap) {
        DimDev Dt = Di;
        int64_t q = dim_lower_bound(Dt, T_f);
        kap[n] = q < Di.n - 1 ? q : Di.n - 1;
      }
      if (var) var[0] = 0;
      if (lag) lag[0] = 0;
      continue;
    }
    const double x = Dj.t[n];
    int64_t q = dim_lower_bound(Di, x);
    if (kap) kap[n] = q - 1;
    if (n < P.H && (var || lag)) {
      while (q < Di.n && Di.t[q] == x) ++q;   // 'right' side: step over ties
      if (var) var[n + 1] = q;
      if (lag) lag[n + 1] = Di.n - q;
    }
  }
}

static int compute_lag_counts(aslsd_ctx* ctx, std::vector<int64_t>& H) {
  const int d = ctx->d;
  for (int i = 0; i < d; ++i)
    if (ctx->n_events[i] == 0)
      return aslsd_fail(ctx, ASLSD_ERR_ARG, "index maps: an event type has no events (IndexError upstream)");
  int64_t* dH = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&dH, sizeof(int64_t) * d * d));
  lag_count_kernel<<<(d * d + 127) / 128, 128, 0, ctx->stream>>>((const DimDev*)ctx->d_dims, d, dH);
  H.resize(d * d);
  cudaError_t e = cudaMemcpyAsync(H.data(), dH, sizeof(int64_t) * d * d, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dH);
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
      ktot += ctx->n_events[j] + 1;
      vtot += (i == j) ? ctx->n_events[i] : H[i * d + j] + 1;
      nmax = std::max(nmax, ctx->n_events[j] + 1);
    }
  PairDev* dP = nullptr;
  int64_t *dK = nullptr, *dV = nullptr, *dL = nullptr;
  cudaError_t e = cudaMalloc(&dP, sizeof(PairDev) * d * d);
  if (e == cudaSuccess && kappa_out) e = cudaMalloc(&dK, sizeof(int64_t) * ktot);
  if (e == cudaSuccess && varpi_out) e = cudaMalloc(&dV, sizeof(int64_t) * vtot);
  if (e == cudaSuccess && lag_sizes_out) e = cudaMalloc(&dL, sizeof(int64_t) * vtot);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(dP, pairs.data(), sizeof(PairDev) * d * d, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    const int threads = 256;
    int64_t bx = (nmax + threads - 1) / threads;
    // grid.x: enough blocks to cover the longest query array, capped at a few
    // waves of the 148 SMs (grid-stride loop beyond that)
    dim3 grid((unsigned)std::min<int64_t>(bx, 148 * 32), (unsigned)(d * d));
    cudaEventRecord(ctx->ev0, ctx->stream);
    index_maps_kernel<<<grid, threads, 0, ctx->stream>>>((const DimDev*)ctx->d_dims, dP, d, ctx->T_f, dK, dV,
                                                         dL);
    cudaEventRecord(ctx->ev1, ctx->stream);
    e = cudaGetLastError();
  }
  for (int p = 0; p < d * d && e == cudaSuccess; ++p) {
    int i = p / d, j = p % d;
    int64_t klen = ctx->n_events[j] + 1;
    int64_t vlen = (i == j) ? ctx->n_events[i] : H[p] + 1;
    if (kappa_out && kappa_out[p])
      e = cudaMemcpyAsync(kappa_out[p], dK + pairs[p].kappa_off, sizeof(int64_t) * klen,
                          cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && varpi_out && varpi_out[p])
      e = cudaMemcpyAsync(varpi_out[p], dV + pairs[p].varpi_off, sizeof(int64_t) * vlen,
                          cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && lag_sizes_out && lag_sizes_out[p])
      e = cudaMemcpyAsync(lag_sizes_out[p], dL + pairs[p].varpi_off, sizeof(int64_t) * vlen,
                          cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms = ms;
    ctx->last_launches = 1;
  }
  cudaFree(dP);
  cudaFree(dK);
  cudaFree(dV);
  cudaFree(dL);
  CUDA_TRY(ctx, e);
  return ASLSD_OK;
}
