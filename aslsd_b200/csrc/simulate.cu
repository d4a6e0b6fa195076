// Cluster (branching) simulation of a linear multivariate Hawkes process on the device.
// Replaces the generation loop of MHP.simulate (models/hawkes/linear/mhp.py:733-854: immigrants
// :790-803, offspring generation by generation :809-845, final sort :848-853) as the source of
// synthetic event streams.  Same law, not the same random numbers: the reference draws from a
// numpy Generator on the host; here every event owns a counter-based Philox4x32-10 stream
// (counter = its slot in the event pool, key = seed), so the result depends on the seed alone.
//
//   immigrants   N_i ~ Poisson(mu_i T_f) (host, transformed rejection), times uniform on [0, T_f)
//   offspring    an event of type j has K ~ Poisson(Lambda_j) children, Lambda_j = sum_{i,r} omega_ijr;
//                each child picks its (type i, basis r) with probability omega_ijr / Lambda_j
//                (splitting of a Poisson variable: the counts per (i, r) are independent
//                Poisson(omega_ijr), as in the reference's loop over kernels and bases) and its lag
//                from the basis' shape: exponential(beta) or N(delta, beta); children with a
//                non-positive lag or beyond T_f are dropped (:826-838)
//   generations  until no event is born; then one sort by (type, time)
//
// HBM-bound integer / byte work around a few transcendental draws; the pool holds (time, type)
// for all generations.  The two radix sorts and the prefix sum use CUB (toolkit header library).
#include "common.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <limits>

// ---- Philox4x32-10 (Salmon et al. 2011), one call = 128 random bits ---------------------------------
struct Philox {
  uint32_t key[2];
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    const uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  __host__ __device__ inline void operator()(uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

// uniform on (0, 1): 53 bits, never 0
__host__ __device__ inline double u01(uint32_t a, uint32_t b) {
  const uint64_t x = (((uint64_t)a << 32) | b) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

// stream of uniforms of one event: slot -> counter, `purpose` separates the uses of a slot
struct EventRng {
  Philox ph;
  uint64_t slot, purpose;
  uint32_t buf[4];
  int left, block;
  __host__ __device__ EventRng(const Philox& p, uint64_t s, uint64_t use) : ph(p), slot(s), purpose(use), left(0), block(0) {}
  __host__ __device__ inline double next() {
    if (left == 0) {
      ph(slot, (purpose << 32) | (uint64_t)block, buf);
      ++block;
      left = 2;
    }
    --left;
    return u01(buf[2 * left], buf[2 * left + 1]);
  }
};

// Poisson by inversion (small means: the branching sums of a stable process are below one)
template <class R>
__host__ __device__ inline int poisson_small(double lam, R& rng) {
  if (!(lam > 0.)) return 0;
  double p = exp(-lam), cdf = p;
  const double u = rng.next();
  int k = 0;
  while (u > cdf && k < 4096) {
    ++k;
    p *= lam / (double)k;
    cdf += p;
  }
  return k;
}

// Poisson of a large mean: transformed rejection (Hoermann 1993, PTRS), host only
static int64_t poisson_large(double lam, EventRng& rng) {
  if (lam < 30.) {
    int64_t k = 0;
    // sum of small pieces keeps the inversion's exp(-lam) away from underflow
    for (int q = 0; q < 8; ++q) k += poisson_small(lam / 8., rng);
    return k;
  }
  const double slam = std::sqrt(lam), loglam = std::log(lam);
  const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.);
  for (;;) {
    const double U = rng.next() - 0.5, V = rng.next();
    const double us = 0.5 - std::fabs(U);
    const double kf = std::floor((2. * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return (int64_t)kf;
    if (kf < 0. || (us < 0.013 && V > us)) continue;
    if (std::log(V) + std::log(inv_alpha) - std::log(a / (us * us) + b) <= -lam + kf * loglam - std::lgamma(kf + 1.))
      return (int64_t)kf;
  }
}

struct SimBasis {
  int32_t child_type, family;   // family: ASLSD_BASIS_*
  double beta, delta;
};

// spawn law of one parent type: total mean and the cumulative table over its (child type, basis) pairs
struct SimLaw {
  double lambda;
  int32_t first, count;   // entries of cdf / basis tables
};

__global__ void sim_immigrants_kernel(double* __restrict__ t, uint8_t* __restrict__ type, int64_t first, int64_t n,
                                      int dim, double T_f, Philox ph) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  EventRng rng(ph, (uint64_t)(first + x), 1);
  t[first + x] = rng.next() * T_f;
  type[first + x] = (uint8_t)dim;
}

__global__ void sim_count_kernel(const double* __restrict__ t, const uint8_t* __restrict__ type, int64_t first,
                                 int64_t n, const SimLaw* __restrict__ laws, int32_t* __restrict__ counts, Philox ph) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const double tp = t[first + x];
  int k = 0;
  if (tp < 1e300) {   // dropped children are parked at +inf
    EventRng rng(ph, (uint64_t)(first + x), 2);
    k = poisson_small(laws[type[first + x]].lambda, rng);
  }
  counts[x] = k;
}

__global__ void sim_fill_kernel(double* __restrict__ t, uint8_t* __restrict__ type, int64_t first, int64_t n,
                                const SimLaw* __restrict__ laws, const double* __restrict__ cdf,
                                const SimBasis* __restrict__ bases, const int32_t* __restrict__ counts,
                                const int64_t* __restrict__ offsets, int64_t child_first, double T_f, Philox ph) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const int k = counts[x];
  if (k == 0) return;
  const double tp = t[first + x];
  const SimLaw L = laws[type[first + x]];
  EventRng rng(ph, (uint64_t)(first + x), 3);
  for (int c = 0; c < k; ++c) {
    // (child type, basis) of this child: inverse of the cumulative table
    const double u = rng.next() * L.lambda;
    int lo = 0, hi = L.count - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[L.first + mid] < u) lo = mid + 1; else hi = mid;
    }
    const SimBasis B = bases[L.first + lo];
    double lag;
    if (B.family == ASLSD_BASIS_EXPONENTIAL) {
      lag = -log(rng.next()) / B.beta;
    } else {
      // N(delta, beta) by Box-Muller (one of the two variates)
      const double r = sqrt(-2. * log(rng.next())), a = 6.283185307179586477 * rng.next();
      lag = B.delta + B.beta * r * cos(a);
    }
    const double tc = tp + lag;
    const int64_t slot = child_first + offsets[x] + c;
    t[slot] = (lag > 0. && tc < T_f) ? tc : 1.7e308;
    type[slot] = (uint8_t)B.child_type;
  }
}

__global__ void sim_keys_kernel(const double* __restrict__ t, uint64_t* __restrict__ keys, int64_t n) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x < n) keys[x] = (uint64_t)__double_as_longlong(t[x]);   // non-negative doubles: bit order = value order
}

// after the sort by (type, time): where every type starts and how many of its events are real
__global__ void sim_bounds_kernel(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ type, int64_t n,
                                  int64_t* __restrict__ first, int64_t* __restrict__ count) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const int ty = type[x];
  const bool real = __longlong_as_double((long long)keys[x]) < 1e300;
  if (x == 0 || type[x - 1] != ty) first[ty] = x;
  const bool next_real = x + 1 < n && type[x + 1] == ty && __longlong_as_double((long long)keys[x + 1]) < 1e300;
  if (real && !next_real) count[ty] = x + 1;   // end of the real events of this type (absolute)
}

struct SimState {
  double* t = nullptr;
  uint8_t* type = nullptr;
  uint64_t *keys_a = nullptr, *keys_b = nullptr;
  uint8_t* type_b = nullptr;
  int32_t* counts = nullptr;
  int64_t* offsets = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  int64_t cap = 0;
  int d = 0;
  std::vector<int64_t> first, count;   // per type, in the sorted key array
  bool sorted_in_a = true;
};

static void sim_free(SimState* S) {
  if (!S) return;
  cudaFree(S->t); cudaFree(S->type); cudaFree(S->keys_a); cudaFree(S->keys_b); cudaFree(S->type_b);
  cudaFree(S->counts); cudaFree(S->offsets); cudaFree(S->tmp);
  delete S;
}
static void sim_free_void(void* p) { sim_free((SimState*)p); }

#define SIM_TRY(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      sim_free(S);                                                                         \
      ctx->sim_state = nullptr;                                                            \
      return aslsd_fail(ctx, ASLSD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    }                                                                                      \
  } while (0)

// kernels: d x d kernels, row-major [i][j] = phi_ij (child type i, parent type j), every basis with
// all of its variables in `fixed` (omega, beta[, delta]); mu: d baselines.  The events stay on the
// device until aslsd_simulate_fetch; n_events receives the counts per type.
extern "C" int aslsd_simulate(aslsd_ctx* ctx, int d, const int32_t* n_basis, const aslsd_basis* bases,
                              const double* mu, double T_f, uint64_t seed, int64_t* n_events) {
  if (!ctx || d <= 0 || d > 255 || !n_basis || !bases || !mu || !n_events || !(T_f > 0.))
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate: bad arguments");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // ---- spawn laws per parent type ---------------------------------------------------------------
  std::vector<SimLaw> laws(d);
  std::vector<double> cdf;
  std::vector<SimBasis> sb;
  std::vector<double> A((size_t)d * d, 0.);   // branching matrix A[i][j] = sum_r omega_ijr
  {
    std::vector<int64_t> kfirst((size_t)d * d + 1, 0);
    for (int q = 0; q < d * d; ++q) {
      if (n_basis[q] < 0) return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate: bad basis count");
      kfirst[q + 1] = kfirst[q] + n_basis[q];
    }
    for (int j = 0; j < d; ++j) {
      laws[j].first = (int32_t)cdf.size();
      double run = 0.;
      for (int i = 0; i < d; ++i)
        for (int r = 0; r < n_basis[i * d + j]; ++r) {
          const aslsd_basis& b = bases[kfirst[i * d + j] + r];
          if (b.type != ASLSD_BASIS_EXPONENTIAL && b.type != ASLSD_BASIS_GAUSSIAN)
            return aslsd_fail(ctx, ASLSD_ERR_UNSUPPORTED, "simulate: basis kernel outside the engine's scope");
          const double omega = b.fixed[0], beta = b.fixed[1];
          if (!(omega >= 0.) || !(beta > 0.)) return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate: omega >= 0, beta > 0");
          if (omega == 0.) continue;
          run += omega;
          A[(size_t)i * d + j] += omega;
          cdf.push_back(run);
          sb.push_back(SimBasis{i, b.type, beta, b.type == ASLSD_BASIS_GAUSSIAN ? b.fixed[2] : 0.});
        }
      laws[j].count = (int32_t)cdf.size() - laws[j].first;
      laws[j].lambda = run;
    }
  }
  // expected counts eta T_f with eta = mu + A eta (a divergent iteration = an unstable process)
  std::vector<double> eta(mu, mu + d);
  for (int it = 0; it < 500; ++it) {
    std::vector<double> nxt(mu, mu + d);
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) nxt[i] += A[(size_t)i * d + j] * eta[j];
    eta.swap(nxt);
  }
  double expected = 0.;
  for (int i = 0; i < d; ++i) {
    if (!(mu[i] >= 0.) || !std::isfinite(eta[i]) || eta[i] > 1e6 * (mu[i] + 1.))
      return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate: the branching matrix is not stable (spectral radius >= 1)");
    expected += eta[i] * T_f;
  }
  if (expected > 1.5e9) return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate: more than ~10^9 events expected");
  if (ctx->sim_state) { sim_free((SimState*)ctx->sim_state); ctx->sim_state = nullptr; }
  SimState* S = new SimState();
  ctx->sim_state = S;
  ctx->sim_state_free = sim_free_void;
  S->d = d;
  // dropped children occupy slots too: ~(1 + share dropped) x the expected count, plus fluctuations
  S->cap = (int64_t)(1.35 * expected + 8. * std::sqrt(expected + 1.) + 4096.);
  SIM_TRY(cudaMalloc((void**)&S->t, sizeof(double) * S->cap));
  SIM_TRY(cudaMalloc((void**)&S->type, S->cap));
  SIM_TRY(cudaMalloc((void**)&S->counts, sizeof(int32_t) * S->cap));
  SIM_TRY(cudaMalloc((void**)&S->offsets, sizeof(int64_t) * S->cap));
  SimLaw* d_laws = nullptr;
  double* d_cdf = nullptr;
  SimBasis* d_sb = nullptr;
  SIM_TRY(cudaMalloc((void**)&d_laws, sizeof(SimLaw) * d));
  SIM_TRY(cudaMalloc((void**)&d_cdf, sizeof(double) * std::max<size_t>(1, cdf.size())));
  SIM_TRY(cudaMalloc((void**)&d_sb, sizeof(SimBasis) * std::max<size_t>(1, sb.size())));
  cudaStream_t st = ctx->stream;
  cudaMemcpyAsync(d_laws, laws.data(), sizeof(SimLaw) * d, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_cdf, cdf.data(), sizeof(double) * cdf.size(), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_sb, sb.data(), sizeof(SimBasis) * sb.size(), cudaMemcpyHostToDevice, st);
  Philox ph;
  ph.key[0] = (uint32_t)seed;
  ph.key[1] = (uint32_t)(seed >> 32);
  auto cleanup = [&]() { cudaFree(d_laws); cudaFree(d_cdf); cudaFree(d_sb); };
  // ---- immigrants -------------------------------------------------------------------------------
  int64_t n_pool = 0;
  {
    EventRng host_rng(ph, 0xFFFFFFFFFFFFull, 0);
    for (int i = 0; i < d; ++i) {
      const int64_t n0 = poisson_large(mu[i] * T_f, host_rng);
      if (n_pool + n0 > S->cap) { cleanup(); sim_free(S); ctx->sim_state = nullptr; return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate: pool overflow"); }
      if (n0 > 0)
        sim_immigrants_kernel<<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(S->t, S->type, n_pool, n0, i, T_f, ph);
      n_pool += n0;
    }
  }
  // ---- generations ------------------------------------------------------------------------------
  int64_t gen_first = 0, gen_n = n_pool;
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, S->counts, S->offsets, (int)std::min<int64_t>(S->cap, 2147483647), st);
  cudaError_t e = cudaMalloc(&S->tmp, scan_bytes + 64);
  S->tmp_bytes = scan_bytes + 64;
  int generations = 0;
  while (e == cudaSuccess && gen_n > 0 && generations < 100000) {
    const unsigned blocks = (unsigned)((gen_n + 255) / 256);
    sim_count_kernel<<<blocks, 256, 0, st>>>(S->t, S->type, gen_first, gen_n, d_laws, S->counts, ph);
    size_t bytes = S->tmp_bytes;
    e = cub::DeviceScan::ExclusiveSum(S->tmp, bytes, S->counts, S->offsets, (int)gen_n, st);
    if (e != cudaSuccess) break;
    int64_t last_off = 0;
    int32_t last_cnt = 0;
    cudaMemcpyAsync(&last_off, S->offsets + gen_n - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&last_cnt, S->counts + gen_n - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) break;
    const int64_t born = last_off + last_cnt;
    if (n_pool + born > S->cap) {
      cleanup(); sim_free(S); ctx->sim_state = nullptr;
      return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate: pool overflow (raise T_f gradually or report this seed)");
    }
    if (born > 0)
      sim_fill_kernel<<<blocks, 256, 0, st>>>(S->t, S->type, gen_first, gen_n, d_laws, d_cdf, d_sb, S->counts,
                                             S->offsets, n_pool, T_f, ph);
    gen_first = n_pool;
    gen_n = born;
    n_pool += born;
    ++generations;
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { cleanup(); SIM_TRY(e); }
  // ---- sort by (type, time): a stable sort by time, then a stable sort by the type byte -------------
  const int64_t n = n_pool;
  S->first.assign(d, 0);
  S->count.assign(d, 0);
  if (n > 0) {
    if (n > 2147483647LL) { cleanup(); sim_free(S); ctx->sim_state = nullptr; return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate: too many events"); }
    e = cudaMalloc((void**)&S->keys_a, sizeof(uint64_t) * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&S->keys_b, sizeof(uint64_t) * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&S->type_b, n);
    if (e != cudaSuccess) { cleanup(); SIM_TRY(e); }
    sim_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S->t, S->keys_a, n);
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, S->keys_a, S->keys_b, S->type, S->type_b, (int)n, 0, 64, st);
    cub::DeviceRadixSort::SortPairs(nullptr, b2, S->type_b, S->type, S->keys_b, S->keys_a, (int)n, 0, 8, st);
    const size_t need = std::max(b1, b2);
    if (need > S->tmp_bytes) {
      cudaFree(S->tmp);
      S->tmp = nullptr;
      e = cudaMalloc(&S->tmp, need);
      S->tmp_bytes = need;
      if (e != cudaSuccess) { cleanup(); SIM_TRY(e); }
    }
    size_t bytes = S->tmp_bytes;
    e = cub::DeviceRadixSort::SortPairs(S->tmp, bytes, S->keys_a, S->keys_b, S->type, S->type_b, (int)n, 0, 64, st);
    bytes = S->tmp_bytes;
    if (e == cudaSuccess)
      e = cub::DeviceRadixSort::SortPairs(S->tmp, bytes, S->type_b, S->type, S->keys_b, S->keys_a, (int)n, 0, 8, st);
    if (e != cudaSuccess) { cleanup(); SIM_TRY(e); }
    // sorted: keys_a (time bits), type
    int64_t* d_bounds = nullptr;
    e = cudaMalloc((void**)&d_bounds, sizeof(int64_t) * 2 * d);
    if (e != cudaSuccess) { cleanup(); SIM_TRY(e); }
    cudaMemsetAsync(d_bounds, 0, sizeof(int64_t) * 2 * d, st);
    sim_bounds_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S->keys_a, S->type, n, d_bounds, d_bounds + d);
    std::vector<int64_t> hb(2 * d);
    cudaMemcpyAsync(hb.data(), d_bounds, sizeof(int64_t) * 2 * d, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    cudaFree(d_bounds);
    if (e != cudaSuccess) { cleanup(); SIM_TRY(e); }
    for (int i = 0; i < d; ++i) {
      S->first[i] = hb[i];
      S->count[i] = hb[d + i] > hb[i] ? hb[d + i] - hb[i] : 0;
    }
  }
  cleanup();
  for (int i = 0; i < d; ++i) n_events[i] = S->count[i];
  ctx->last_launches = generations;
  return ASLSD_OK;
}

// Copy the simulated event times of every type to the caller's arrays (n_events[i] doubles each).
extern "C" int aslsd_simulate_fetch(aslsd_ctx* ctx, double* const* times) {
  if (!ctx || !times) return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate_fetch: bad arguments");
  SimState* S = (SimState*)ctx->sim_state;
  if (!S) return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate_fetch: no simulation on this context");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  for (int i = 0; i < S->d; ++i)
    if (S->count[i] > 0) {
      if (!times[i]) return aslsd_fail(ctx, ASLSD_ERR_ARG, "simulate_fetch: NULL output array");
      CUDA_TRY(ctx, cudaMemcpyAsync(times[i], S->keys_a + S->first[i], sizeof(double) * S->count[i],
                                    cudaMemcpyDeviceToHost, ctx->stream));
    }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  sim_free(S);
  ctx->sim_state = nullptr;
  return ASLSD_OK;
}
