// Lag-tiled pairwise kernels: the O(N * window) double sums of the exact LSE.
//
// One CTA = one work unit = (item, block of target events).  An item is one
// basis pair (upsilon) or one basis (phi) of one named term.  Per tile of TM
// targets the CTA
//   1. finds, per target m, the source range [lo(m), hi(m)) whose lags fall in
//      the item's window (bucket-table lower_bound, common.cuh),
//   2. stages that event-time window in shared memory with one TMA bulk copy
//      (cp.async.bulk + mbarrier),
//   3. flattens the (m, n) band into a dense pair index so that all 32 lanes of
//      every warp evaluate a pair (fp64 pipe is the bound, not HBM),
//   4. accumulates the raw components in registers; one shuffle + shared
//      reduction per work unit writes a partial, summed in fixed order by
//      finalize_kernel (deterministic).
#pragma once

#include "basis_math.cuh"

#include "pair_tiles.h"

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// locate the work unit's item: last item with wu_first <= wu
__device__ __forceinline__ int find_item(const ItemDev* __restrict__ items, int n_items, int64_t wu) {
  int lo = 0, hi = n_items;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (items[mid].wu_first <= wu) lo = mid; else hi = mid;
  }
  return lo;
}

template <int NACC>
__device__ __forceinline__ void block_reduce_store(double* acc, unsigned long long cnt, double* s_red,
                                                   double* __restrict__ out, unsigned long long* cnt_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < NACC; ++c) {
    double v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp * ASLSD_MAX_COMP + c] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  __shared__ unsigned long long s_cnt[PAIR_NT / 32];
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x < ASLSD_MAX_COMP) {
    double v = 0.;
    if (threadIdx.x < NACC)
      for (int w = 0; w < PAIR_NT / 32; ++w) v += s_red[w * ASLSD_MAX_COMP + threadIdx.x];
    out[threadIdx.x] = v;
  }
  if (threadIdx.x == 32) {
    unsigned long long c = 0;
    for (int w = 0; w < PAIR_NT / 32; ++w) c += s_cnt[w];
    *cnt_out = c;
  }
}

// FAMILY: ITEM_UPS_EXP, ITEM_UPS_GAUSS, ITEM_UPS_EG, ITEM_UPS_GE, ITEM_PHI_EXP, ITEM_PHI_GAUSS
template <int FAMILY>
__global__ void __launch_bounds__(PAIR_NT, 2)
pair_kernel(const DimDev* __restrict__ dims, const ItemDev* __restrict__ items, int n_items, double T_f,
            double* __restrict__ partials, unsigned long long* __restrict__ counts, int64_t wu_base,
            const double* __restrict__ weights) {
  __shared__ __align__(16) double s_src[PAIR_SMAX + 2];
  __shared__ double s_t[PAIR_TM];
  __shared__ int64_t s_lo[PAIR_TM];       // lo(m): first in-window source of target m
  __shared__ int64_t s_pre[PAIR_TM + 1];  // exclusive prefix of pair counts
  __shared__ int64_t s_warp[PAIR_NT / 32];
  __shared__ double s_red[(PAIR_NT / 32) * ASLSD_MAX_COMP];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int64_t s_bounds[2];

  const int tid = threadIdx.x;
  const int64_t wu = (int64_t)blockIdx.x + wu_base;
  const int it_ix = find_item(items, n_items, wu);
  const ItemDev it = items[it_ix];
  const int chunk = (int)(wu - it.wu_first);
  const DimDev Dt = dims[it.tgt_dim], Ds = dims[it.src_dim];
  if (*Dt.bad) return;   // unsorted upload (aslsd_set_events): the verdict reaches the host later
  const int64_t m_begin = it.m_lo + (int64_t)chunk * it.tpw;
  const int64_t m_end = min(it.m_hi, m_begin + it.tpw);
  const int mask = it.mask;
  const bool same = it.same != 0;
  const bool weighted = it.wa_off >= 0 || it.wb_off >= 0;

  UpsGaussConst KG;
  UpsExpConst KE;
  PhiConst KP;
  UpsMixConst KM;
  if (FAMILY == ITEM_UPS_GAUSS) KG.init(it);
  if (FAMILY == ITEM_UPS_EXP) KE.init(it);
  if (FAMILY == ITEM_UPS_EG) KM.init_eg(it);
  if (FAMILY == ITEM_UPS_GE) KM.init_ge(it);
  if (FAMILY == ITEM_PHI_EXP || FAMILY == ITEM_PHI_GAUSS) KP.init(it);

  double acc[ASLSD_MAX_COMP];
#pragma unroll
  for (int c = 0; c < ASLSD_MAX_COMP; ++c) acc[c] = 0.;
  unsigned long long n_eval = 0;

  if (tid == 0) mbar_init(&s_bar, 1);
  fence_proxy_async();
  __syncthreads();
  uint32_t parity = 0;

  for (int64_t m0 = m_begin; m0 < m_end; m0 += PAIR_TM) {
    const int nt = (int)min((int64_t)PAIR_TM, m_end - m0);
    // ---- 1. per-target source ranges -------------------------------------------------
    int64_t my_cnt[PAIR_TM / PAIR_NT];
    int64_t my_sum = 0;
#pragma unroll
    for (int u = 0; u < PAIR_TM / PAIR_NT; ++u) {
      const int x = tid * (PAIR_TM / PAIR_NT) + u;
      int64_t cnt = 0;
      if (x < nt) {
        const double tm = Dt.t[m0 + x];
        s_t[x] = tm;
        int64_t hi = same ? (m0 + x) : dim_lower_bound(Ds, tm);
        if (it.s_lo > 0.) hi = min(hi, dim_upper_bound(Ds, tm - it.s_lo));
        int64_t lo = dim_lower_bound(Ds, tm - it.s_hi);
        lo = min(lo, hi);
        s_lo[x] = lo;
        cnt = hi - lo;
        // lo(m) and hi(m) are monotone in m: the tile needs sources [lo(first), hi(last))
        if (x == 0) s_bounds[0] = lo;
        if (x == nt - 1) s_bounds[1] = hi;
      }
      my_cnt[u] = cnt;
      my_sum += cnt;
    }
    // ---- 2. exclusive scan of the pair counts ------------------------------------------
    int64_t incl = my_sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    int64_t base = 0;
    for (int w = 0; w < (tid >> 5); ++w) base += s_warp[w];
    int64_t run = base + incl - my_sum;
#pragma unroll
    for (int u = 0; u < PAIR_TM / PAIR_NT; ++u) {
      const int x = tid * (PAIR_TM / PAIR_NT) + u;
      if (x < nt) s_pre[x] = run;
      run += my_cnt[u];
    }
    if (tid == PAIR_NT - 1) s_pre[nt] = run;   // total number of pairs of the tile
    __syncthreads();
    const int64_t total = s_pre[nt];
    const int64_t src_first = s_bounds[0], src_end = s_bounds[1];
    if (total == 0) { __syncthreads(); continue; }
    // ---- 3. stage the source window ------------------------------------------------------
    const int64_t a0 = src_first & ~(int64_t)1;
    const int64_t n_stage = ((src_end - a0) + 1) & ~(int64_t)1;
    const bool staged = n_stage <= PAIR_SMAX;
    if (staged) {
      if (tid == 0) {
        mbar_expect_tx(&s_bar, (uint32_t)(n_stage * sizeof(double)));
        tma_load_1d(s_src, Ds.t + a0, (uint32_t)(n_stage * sizeof(double)), &s_bar);
      }
      mbar_wait(&s_bar, parity);
      parity ^= 1;
    }
    const double* srcp = staged ? (s_src - a0) : Ds.t;
    const bool interior = (T_f - s_t[nt - 1]) >= it.t_int;
    // ---- 4. dense pair loop -------------------------------------------------------------------
    for (int64_t p = tid; p < total; p += PAIR_NT) {
      int lo = 0, hi = nt;   // last x with s_pre[x] <= p
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_pre[mid] <= p) lo = mid; else hi = mid;
      }
      const double tm = s_t[lo];
      const int64_t n = s_lo[lo] + (p - s_pre[lo]);
      const double s = tm - srcp[n];
      // per-event weights of a marked model: impact of the target times impact of the source
      // (mtlh_exact_estimator.py:838, :896-897); 1 for unmarked models
      double w = 1.;
      if (weighted) {
        if (it.wa_off >= 0) w = weights[it.wa_off + m0 + lo];
        if (it.wb_off >= 0) w *= weights[it.wb_off + n];
      }
      double v[ASLSD_MAX_COMP];
#pragma unroll
      for (int c = 0; c < ASLSD_MAX_COMP; ++c) v[c] = 0.;
      if (FAMILY == ITEM_UPS_GAUSS) {
        if (interior) eval_ups_gauss<true>(KG, mask, 0., s, v);
        else eval_ups_gauss<false>(KG, mask, T_f - tm, s, v);
      } else if (FAMILY == ITEM_UPS_EXP) {
        if (interior) eval_ups_exp<true>(KE, mask, 0., s, v);
        else eval_ups_exp<false>(KE, mask, T_f - tm, s, v);
      } else if (FAMILY == ITEM_UPS_EG) {
        if (interior) eval_ups_eg<true>(KM, mask, 0., s, v);
        else eval_ups_eg<false>(KM, mask, T_f - tm, s, v);
      } else if (FAMILY == ITEM_UPS_GE) {
        if (interior) eval_ups_ge<true>(KM, mask, 0., s, v);
        else eval_ups_ge<false>(KM, mask, T_f - tm, s, v);
      } else if (FAMILY == ITEM_PHI_EXP) {
        eval_phi_exp(KP, mask, s, v);
      } else {
        eval_phi_gauss(KP, mask, s, v);
      }
#pragma unroll
      for (int c = 0; c < ASLSD_MAX_COMP; ++c) acc[c] = fma(w, v[c], acc[c]);
    }
    n_eval += (unsigned long long)((total - tid + PAIR_NT - 1) / PAIR_NT) * (tid < total ? 1 : 0);
    fence_proxy_async();   // generic reads of s_src before the next tile's async write
    __syncthreads();
  }
  block_reduce_store<ASLSD_MAX_COMP>(acc, n_eval, s_red, partials + wu * ASLSD_MAX_COMP, counts + wu);
}

// Single-index sums: psi_i (one evaluation per target) and upsilon at lag zero.
// FAMILY: ITEM_PSI_EXP, ITEM_PSI_GAUSS, ITEM_UPS0_EXP, ITEM_UPS0_GAUSS, ITEM_UPS0_EG, ITEM_UPS0_GE
template <int FAMILY>
__global__ void __launch_bounds__(PAIR_NT, 2)
single_kernel(const DimDev* __restrict__ dims, const ItemDev* __restrict__ items, int n_items, double T_f,
              double* __restrict__ partials, unsigned long long* __restrict__ counts, int64_t wu_base,
              const double* __restrict__ weights) {
  __shared__ double s_red[(PAIR_NT / 32) * ASLSD_MAX_COMP];
  const int tid = threadIdx.x;
  const int64_t wu = (int64_t)blockIdx.x + wu_base;
  const int it_ix = find_item(items, n_items, wu);
  const ItemDev it = items[it_ix];
  const int chunk = (int)(wu - it.wu_first);
  const DimDev Dt = dims[it.tgt_dim];
  const int64_t m_begin = it.m_lo + (int64_t)chunk * it.tpw;
  const int64_t m_end = min(it.m_hi, m_begin + it.tpw);
  const int mask = it.mask;
  UpsGaussConst KG;
  UpsExpConst KE;
  PhiConst KP;
  UpsMixConst KM;
  if (FAMILY == ITEM_UPS0_GAUSS) KG.init(it);
  if (FAMILY == ITEM_UPS0_EXP) KE.init(it);
  if (FAMILY == ITEM_UPS0_EG) KM.init_eg(it);
  if (FAMILY == ITEM_UPS0_GE) KM.init_ge(it);
  if (FAMILY == ITEM_PSI_EXP || FAMILY == ITEM_PSI_GAUSS) KP.init(it);
  double acc[ASLSD_MAX_COMP];
#pragma unroll
  for (int c = 0; c < ASLSD_MAX_COMP; ++c) acc[c] = 0.;
  unsigned long long n_eval = 0;
  for (int64_t m = m_begin + tid; m < m_end; m += PAIR_NT) {
    const double t = T_f - Dt.t[m];
    const bool interior = t >= it.t_int;
    const double w = it.wa_off >= 0 ? weights[it.wa_off + m] : 1.;   // marked models
    double v[ASLSD_MAX_COMP];
#pragma unroll
    for (int c = 0; c < ASLSD_MAX_COMP; ++c) v[c] = 0.;
    if (FAMILY == ITEM_UPS0_GAUSS) {
      if (interior) eval_ups_gauss<true>(KG, mask, 0., 0., v);
      else eval_ups_gauss<false>(KG, mask, t, 0., v);
    } else if (FAMILY == ITEM_UPS0_EXP) {
      if (interior) eval_ups_exp<true>(KE, mask, 0., 0., v);
      else eval_ups_exp<false>(KE, mask, t, 0., v);
    } else if (FAMILY == ITEM_UPS0_EG) {
      if (interior) eval_ups_eg<true>(KM, mask, 0., 0., v);
      else eval_ups_eg<false>(KM, mask, t, 0., v);
    } else if (FAMILY == ITEM_UPS0_GE) {
      if (interior) eval_ups_ge<true>(KM, mask, 0., 0., v);
      else eval_ups_ge<false>(KM, mask, t, 0., v);
    } else if (FAMILY == ITEM_PSI_EXP) {
      eval_psi_exp(KP, mask, t, v);
    } else {
      eval_psi_gauss(KP, mask, t, v);
    }
#pragma unroll
    for (int c = 0; c < ASLSD_MAX_COMP; ++c) acc[c] = fma(w, v[c], acc[c]);
    ++n_eval;
  }
  if (chunk == 0 && tid == 0 && it.n_int > 0) {
    // interior targets: the closed form no longer depends on T_f - t_m (upper limits at
    // +infinity, exactly what fp64 gives there): one evaluation times their number
    double v[ASLSD_MAX_COMP];
#pragma unroll
    for (int c = 0; c < ASLSD_MAX_COMP; ++c) v[c] = 0.;
    const double t_far = 1e300;
    if (FAMILY == ITEM_UPS0_GAUSS) eval_ups_gauss<true>(KG, mask, 0., 0., v);
    else if (FAMILY == ITEM_UPS0_EXP) eval_ups_exp<true>(KE, mask, 0., 0., v);
    else if (FAMILY == ITEM_UPS0_EG) eval_ups_eg<true>(KM, mask, 0., 0., v);
    else if (FAMILY == ITEM_UPS0_GE) eval_ups_ge<true>(KM, mask, 0., 0., v);
    else if (FAMILY == ITEM_PSI_EXP) eval_psi_exp(KP, mask, t_far, v);
    else eval_psi_gauss(KP, mask, t_far, v);
#pragma unroll
    for (int c = 0; c < ASLSD_MAX_COMP; ++c) acc[c] += (double)it.n_int * v[c];
    ++n_eval;
  }
  block_reduce_store<ASLSD_MAX_COMP>(acc, n_eval, s_red, partials + wu * ASLSD_MAX_COMP, counts + wu);
}

// raw[dst_row(item)][c] += sum over the item's work units, in order.
__global__ void finalize_kernel(const ItemDev* __restrict__ items, int n_items,
                                const double* __restrict__ partials,
                                const unsigned long long* __restrict__ counts, double* __restrict__ raw,
                                unsigned long long* __restrict__ item_counts) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int item = g / (ASLSD_MAX_COMP + 1), c = g % (ASLSD_MAX_COMP + 1);
  if (item >= n_items) return;
  const ItemDev it = items[item];
  if (c < ASLSD_MAX_COMP) {
    double v = 0.;
    for (int w = 0; w < it.n_chunks; ++w) v += partials[(it.wu_first + w) * ASLSD_MAX_COMP + c];
    // every item owns its destination row (raw is zeroed before the launches; the
    // class kernel's finalize adds the interior part of the same rows afterwards)
    raw[it.dst_row * ASLSD_MAX_COMP + c] += v;
  } else {
    unsigned long long v = 0;
    for (int w = 0; w < it.n_chunks; ++w) v += counts[it.wu_first + w];
    item_counts[item] = v;
  }
}

// In-window pair count of a double-sum term (metric unit): sum_m (hi(m) - lo(m)).
__global__ void __launch_bounds__(256)
count_pairs_kernel(const DimDev* __restrict__ dims, int tgt_dim, int src_dim, int same, int64_t m_lo,
                   int64_t m_hi, double s_lo, double s_hi, unsigned long long* __restrict__ out) {
  const DimDev Dt = dims[tgt_dim], Ds = dims[src_dim];
  unsigned long long cnt = 0;
  for (int64_t m = m_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < m_hi;
       m += (int64_t)gridDim.x * blockDim.x) {
    const double tm = Dt.t[m];
    int64_t hi = same ? m : dim_lower_bound(Ds, tm);
    if (s_lo > 0.) hi = min(hi, dim_upper_bound(Ds, tm - s_lo));
    int64_t lo = min(dim_lower_bound(Ds, tm - s_hi), hi);
    cnt += (unsigned long long)(hi - lo);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, cnt);
}
