// Per-event sums over the earlier events: the intensity at the jump times and the kernel part of
// the compensator, the two O(N^2) loops that follow a fit in the reference.
//   kind 0 (phi):  out[i][m] = sum_{n: t^i_n < t^k_m} phi_ki(t^k_m - t^i_n)
//                  MHP.get_intensity_at_jumps, models/hawkes/linear/mhp.py:669-690
//   kind 1 (psi):  out[i][m] = sum_{n: t^i_n < t^k_m} psi_ki(t^k_m - t^i_n)
//                  kernel_part of get_residuals_k_mhp, stats/residual_analysis/goodness_of_fit.py:169-228
// (for i == k the sum runs over n < m by index, as kappa[i][i] does: time_ordering.py:44-47).
// One thread per target event: source window through the bucket table, closed forms of
// basis_math.cuh per basis.  phi is negligible beyond the kernels' effective support (plan.cu
// windows); psi has reached its limit there (Phi(9) = 1 - 1e-19, e^-42 = 6e-19), so the events
// before the window contribute (their number) x psi(+inf).
#include "basis_math.cuh"
#include "plan.h"

#include <cmath>
#include <cstring>

#define PS_MAX_BASES 64

struct PointBasisDev {
  int32_t type, pad;
  double coef;      // omega beta (exponential phi) or omega
  double beta, delta;
  double limit;     // psi(+inf) of the basis (kind 1)
};

struct PointTaskDev {
  int32_t src_dim, same, n_bases, pad;
  double window;
  double limit;     // sum of the bases' psi(+inf)
  PointBasisDev bases[PS_MAX_BASES];
};

__global__ void __launch_bounds__(256)
point_sums_kernel(const DimDev* __restrict__ dims, int tgt_dim, const PointTaskDev* __restrict__ tasks, int kind,
                  double* __restrict__ out) {
  __shared__ PointTaskDev T;
  const PointTaskDev* src = tasks + blockIdx.y;
  for (int x = threadIdx.x; x < (int)(sizeof(PointTaskDev) / sizeof(double)); x += blockDim.x)
    ((double*)&T)[x] = ((const double*)src)[x];
  __syncthreads();
  const DimDev Dt = dims[tgt_dim], Ds = dims[T.src_dim];
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Dt.n) return;
  const double tm = Dt.t[m];
  const int64_t hi = T.same ? m : dim_lower_bound(Ds, tm);
  const int64_t lo = min(dim_lower_bound(Ds, tm - T.window), hi);
  double acc = 0.;
  for (int64_t n = lo; n < hi; ++n) {
    const double s = tm - Ds.t[n];
    double v = 0.;
    for (int r = 0; r < T.n_bases; ++r) {
      const PointBasisDev B = T.bases[r];
      double f;
      if (kind == 0) {
        if (B.type == ASLSD_BASIS_EXPONENTIAL) {
          f = exp(-B.beta * s);
        } else {
          const double z = (s - B.delta) / B.beta;
          f = (1. / (B.beta * sqrt(2. * 3.14159265358979323846))) * exp(-0.5 * z * z);
        }
      } else {
        if (B.type == ASLSD_BASIS_EXPONENTIAL) f = 1. - exp(-B.beta * s);
        else f = normcdf((s - B.delta) / B.beta) - normcdf(-B.delta / B.beta);
      }
      v = fma(B.coef, f, v);
    }
    acc += v;
  }
  if (kind == 1) acc = fma((double)lo, T.limit, acc);
  out[(int64_t)blockIdx.y * Dt.n + m] = acc;
}

extern "C" int aslsd_point_sums(aslsd_ctx* ctx, int k, const double* x_k, int kind, double* out) {
  if (!ctx || !x_k || !out || (kind != 0 && kind != 1))
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "point_sums: bad arguments");
  if (!ctx->has_model) return aslsd_fail(ctx, ASLSD_ERR_ARG, "point_sums: no model set");
  if (!ctx->d_times || (int)ctx->n_events.size() != ctx->d)
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "point_sums: no events set");
  if (int vrc = events_verdict(ctx, true)) return vrc;
  const int d = ctx->d;
  if (k < 0 || k >= d) return aslsd_fail(ctx, ASLSD_ERR_ARG, "point_sums: bad k");
  SumPlan P;
  P.k = k;
  int rc = plan_vars(ctx, k, x_k, P);
  if (rc) return rc;
  std::vector<PointTaskDev> tasks((size_t)d);
  for (int i = 0; i < d; ++i) {
    const KernelHost& K = ctx->kernels[k * d + i];
    if (K.bases.size() > PS_MAX_BASES)
      return aslsd_fail(ctx, ASLSD_ERR_UNSUPPORTED, "point_sums: more than 64 bases in one kernel");
    PointTaskDev& T = tasks[(size_t)i];
    std::memset(&T, 0, sizeof(T));
    T.src_dim = i;
    T.same = (i == k) ? 1 : 0;
    T.n_bases = (int32_t)K.bases.size();
    for (size_t r = 0; r < K.bases.size(); ++r) {
      const Vars& v = P.V[i][r];
      PointBasisDev& B = T.bases[r];
      B.type = K.bases[r].type;
      B.beta = v[1];
      B.delta = v[2];
      if (B.type == ASLSD_BASIS_EXPONENTIAL) {
        B.coef = kind == 0 ? v[0] * v[1] : v[0];
        B.limit = v[0];
        T.window = std::max(T.window, 42.0 / v[1]);
      } else {
        B.coef = v[0];
        B.limit = v[0] * (1. - 0.5 * std::erfc(v[2] / v[1] / std::sqrt(2.)));   // omega Phi(delta / beta)
        T.window = std::max(T.window, v[2] + 9.0 * v[1]);
      }
      T.limit += B.limit;
    }
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int64_t n_k = ctx->n_events[k];
  const size_t out_bytes = (size_t)d * (size_t)std::max<int64_t>(n_k, 1) * sizeof(double);
  const size_t task_bytes = tasks.size() * sizeof(PointTaskDev);
  if ((rc = ensure_device_buffer(ctx, &ctx->d_sample, &ctx->sample_cap, out_bytes + task_bytes + 64))) return rc;
  if ((rc = ensure_pinned_buffer(ctx, std::max(out_bytes, task_bytes)))) return rc;
  char* dp = (char*)ctx->d_sample;
  const size_t o_out = (task_bytes + 63) & ~(size_t)63;
  std::memcpy(ctx->h_pinned, tasks.data(), task_bytes);
  CUDA_TRY(ctx, cudaMemcpyAsync(dp, ctx->h_pinned, task_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  ctx->last_launches = 0;
  if (n_k > 0) {
    const dim3 grid((unsigned)((n_k + 255) / 256), (unsigned)d);
    point_sums_kernel<<<grid, 256, 0, ctx->stream>>>((const DimDev*)ctx->d_dims, k, (const PointTaskDev*)dp, kind,
                                                     (double*)(dp + o_out));
    ctx->last_launches++;
  }
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timing_pending = true;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));     // the staging block is reused for the result
  if (n_k > 0) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pinned, dp + o_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(out, ctx->h_pinned, out_bytes);
  }
  return ASLSD_OK;
}
