// Shared-class pair kernel: interior targets of the Gaussian double sums whose
// raw components are products of two single-variable factors of the lag s,
//
//     G_pq(s)  = P_e(s) * Q_n(s),   P_e(s) = pn_e exp(-((s - dd_e) isig_e)^2 / 2),
//                                   Q_n(s) = 1 - Phi(c0_n + c1_n s),
//     (upsilon of basis pair (p, q): e = class of (delta_q - delta_p, sigma_pq),
//      n = class of the lower integration limit; basis_kernel_gaussian.py:21-34
//      with the upper limit at +infinity, i.e. T_f - t_m beyond the kernels' reach)
//     SW_p(s)  = 2 I0_p * P_e(s)     (own-basis d/d omega, :507-508; constant I0_p)
//     phi_r(s) = P_e(s)              (basis_kernel_gaussian.py:448-452)
//
// With B bases there are B^2 (p, q) pairs but only O(B) distinct P and Q classes
// when the bases share beta / sit on a delta grid (the SBF models of the
// reference, papers/grad_based_estim/experiments/fit_sbgauss1d10r.py:60-63), and
// the same (tgt, src) event-pair stream serves every (p, q): one exp per P
// class and one normcdf per Q class are evaluated per event pair instead of one
// of each per basis pair, and the products are reduced by the warp.
//
// One CTA = one work unit = (class term, block of targets).  Per tile of CK_TM
// targets: window bounds through the bucket table, exclusive scan of the pair
// counts, TMA bulk copy of the source window into shared memory (as in
// pair_kernels.cuh).  Then, per warp and per batch of 32 event pairs:
//   phase 1 (lane = event pair): all P_e(s) and Q_n(s) -> per-warp shared tiles;
//   phase 2 (lane = product):    acc[prod] += sum_l P[l][e(prod)] * Q[l][n(prod)].
// Accumulators live in registers (CK_MAXR per lane), are combined across the
// CTA's warps in shared memory and written as one partial per work unit;
// class_finalize_kernel sums partials in fixed order (deterministic).
#pragma once

#include "basis_math.cuh"
#include "fast_math.cuh"
#include "pair_kernels.cuh"

#define CK_NT 256
#define CK_TM 512
#define CK_SMAX 2048
#define CK_MAXR 8            // products per lane: at most 32 * CK_MAXR products per class term
#define CK_MAX_CLASSES 40    // P classes and Q classes per class term

struct EClassDev {   // P_e(s) = pn exp(-0.5 ((s - dd) isig)^2)
  double dd, isig, pn;
};
struct NClassDev {   // Q_n(s) = 1 - Phi(sqrt(2) (c0 + c1 s)) = erfc(c0 + c1 s) / 2; one != 0: Q = 1
  double c0, c1;
  int32_t one, pad;
};
struct ProdDev {     // product -> classes and destination
  int16_t e, n;
  int32_t dst;       // index into the raw accumulator array: item * ASLSD_MAX_COMP + comp
  double scale;      // constant factor (2 I0_p for SW, 1 otherwise)
};
struct ClassTermDev {
  int32_t tgt_dim, src_dim, same;
  int32_t nE, nN, nProd;
  int32_t e_first, n_first, prod_first;
  int32_t n_chunks;
  int64_t m_lo, m_hi;      // interior target block
  int64_t tpw;             // targets per work unit
  int64_t wu_first;        // first work unit (prefix over class terms)
  int64_t part_first;      // offset of the term's partials: wu-major, nProd doubles each
  double s_lo, s_hi;       // union lag window of the classes
};

__device__ __forceinline__ int find_class_term(const ClassTermDev* __restrict__ terms, int n_terms, int64_t wu) {
  int lo = 0, hi = n_terms;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (terms[mid].wu_first <= wu) lo = mid; else hi = mid;
  }
  return lo;
}

// dynamic shared memory: [source window | target times | per-warp class tiles
// (aliased by the final reduction) | class and product tables | lo | prefix]
__host__ __device__ inline size_t class_smem_bytes(int nE, int nN, int nProd) {
  const int sE = nE | 1, sN = nN | 1;
  size_t b = 0;
  b += (CK_SMAX + 2) * sizeof(double);
  b += CK_TM * sizeof(double);
  size_t tiles = (size_t)(CK_NT / 32) * 32 * (sE + sN) * sizeof(double);
  size_t red = (size_t)(CK_NT / 32) * 32 * CK_MAXR * sizeof(double);
  b += tiles > red ? tiles : red;
  b += nE * sizeof(EClassDev) + nN * sizeof(NClassDev) + nProd * sizeof(ProdDev);
  b += (CK_TM + CK_TM + 1) * sizeof(int32_t);
  return b + 64;
}

__global__ void __launch_bounds__(CK_NT, 2)
class_kernel(const DimDev* __restrict__ dims, const ClassTermDev* __restrict__ terms, int n_terms,
             const EClassDev* __restrict__ ecls, const NClassDev* __restrict__ ncls,
             const ProdDev* __restrict__ prods, double* __restrict__ partials,
             unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int64_t s_bounds[2];
  __shared__ int32_t s_warp[CK_NT / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t wu = blockIdx.x;
  const ClassTermDev T = terms[find_class_term(terms, n_terms, wu)];
  const int chunk = (int)(wu - T.wu_first);
  const DimDev Dt = dims[T.tgt_dim], Ds = dims[T.src_dim];
  const int64_t m_begin = T.m_lo + (int64_t)chunk * T.tpw;
  const int64_t m_end = min(T.m_hi, m_begin + T.tpw);
  const bool same = T.same != 0;
  const int nE = T.nE, nN = T.nN, nProd = T.nProd;
  const int sE = nE | 1, sN = nN | 1;   // odd strides: conflict-free lane-major tiles

  // carve the dynamic shared memory
  double* s_src = (double*)smem_raw;
  double* s_t = s_src + (CK_SMAX + 2);
  double* s_tiles = s_t + CK_TM;
  const size_t tiles_d = (size_t)(CK_NT / 32) * 32 * (sE + sN);
  const size_t red_d = (size_t)(CK_NT / 32) * 32 * CK_MAXR;
  double* after = s_tiles + (tiles_d > red_d ? tiles_d : red_d);
  EClassDev* s_e = (EClassDev*)after;
  NClassDev* s_n = (NClassDev*)(s_e + nE);
  ProdDev* s_p = (ProdDev*)(s_n + nN);
  int32_t* s_lo = (int32_t*)(s_p + nProd);
  int32_t* s_pre = s_lo + CK_TM;
  double* myE = s_tiles + (size_t)warp * 32 * (sE + sN);
  double* myN = myE + 32 * sE;

  for (int x = tid; x < nE; x += CK_NT) s_e[x] = ecls[T.e_first + x];
  for (int x = tid; x < nN; x += CK_NT) s_n[x] = ncls[T.n_first + x];
  for (int x = tid; x < nProd; x += CK_NT) s_p[x] = prods[T.prod_first + x];
  if (tid == 0) mbar_init(&s_bar, 1);
  fence_proxy_async();
  __syncthreads();
  uint32_t parity = 0;

  // this lane's products
  int pe[CK_MAXR], pn[CK_MAXR];
  double acc[CK_MAXR];
#pragma unroll
  for (int r = 0; r < CK_MAXR; ++r) {
    const int q = r * 32 + lane;
    acc[r] = 0.;
    pe[r] = q < nProd ? s_p[q].e : 0;
    pn[r] = q < nProd ? s_p[q].n : 0;
  }
  const int n_rounds = (nProd + 31) >> 5;
  unsigned long long n_pairs = 0;

  for (int64_t m0 = m_begin; m0 < m_end; m0 += CK_TM) {
    const int nt = (int)min((int64_t)CK_TM, m_end - m0);
    // ---- 1. per-target source ranges (CK_TM / CK_NT = 2 targets per thread) ---------------
    int64_t lo_abs[CK_TM / CK_NT];
    int32_t my_cnt[CK_TM / CK_NT];
    int32_t my_sum = 0;
#pragma unroll
    for (int u = 0; u < CK_TM / CK_NT; ++u) {
      const int x = tid * (CK_TM / CK_NT) + u;
      int32_t cnt = 0;
      lo_abs[u] = 0;
      if (x < nt) {
        const double tm = Dt.t[m0 + x];
        s_t[x] = tm;
        int64_t hi = same ? (m0 + x) : dim_lower_bound(Ds, tm);
        if (T.s_lo > 0.) hi = min(hi, dim_upper_bound(Ds, tm - T.s_lo));
        int64_t lo = min(dim_lower_bound(Ds, tm - T.s_hi), hi);
        lo_abs[u] = lo;
        cnt = (int32_t)(hi - lo);
        if (x == 0) s_bounds[0] = lo;
        if (x == nt - 1) s_bounds[1] = hi;
      }
      my_cnt[u] = cnt;
      my_sum += cnt;
    }
    // ---- 2. exclusive scan of the pair counts ---------------------------------------------
    int32_t incl = my_sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int32_t run = incl - my_sum;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
    const int64_t src_first = s_bounds[0], src_end = s_bounds[1];
#pragma unroll
    for (int u = 0; u < CK_TM / CK_NT; ++u) {
      const int x = tid * (CK_TM / CK_NT) + u;
      if (x < nt) {
        s_pre[x] = run;
        s_lo[x] = (int32_t)(lo_abs[u] - src_first);
      }
      run += my_cnt[u];
    }
    if (tid == CK_NT - 1) s_pre[nt] = run;
    __syncthreads();
    const int32_t total = s_pre[nt];
    if (total == 0) { __syncthreads(); continue; }
    // ---- 3. stage the source window -------------------------------------------------------
    const int64_t a0 = src_first & ~(int64_t)1;
    const int64_t n_stage = ((src_end - a0) + 1) & ~(int64_t)1;
    const bool staged = n_stage <= CK_SMAX;
    if (staged) {
      if (tid == 0) {
        mbar_expect_tx(&s_bar, (uint32_t)(n_stage * sizeof(double)));
        tma_load_1d(s_src, Ds.t + a0, (uint32_t)(n_stage * sizeof(double)), &s_bar);
      }
      mbar_wait(&s_bar, parity);
      parity ^= 1;
    }
    const double* srcp = staged ? (s_src + (src_first - a0)) : (Ds.t + src_first);
    // ---- 4. pairs, 32 per warp at a time --------------------------------------------------
    for (int32_t p0 = warp * 32; p0 < total; p0 += CK_NT) {
      const int32_t p = p0 + lane;
      const bool live = p < total;
      double s = 0.;
      if (live) {
        int lo = 0, hi = nt;   // last target x with s_pre[x] <= p
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (s_pre[mid] <= p) lo = mid; else hi = mid;
        }
        s = s_t[lo] - srcp[s_lo[lo] + (p - s_pre[lo])];
      }
      // phase 1: every class at this lane's lag (two independent chains per step)
      double* rowE = myE + lane * sE;
      double* rowN = myN + lane * sN;
      for (int e = 0; e < nE; e += 2) {
        const EClassDev A = s_e[e];
        const EClassDev B = s_e[min(e + 1, nE - 1)];
        const double za = (s - A.dd) * A.isig, zb = (s - B.dd) * B.isig;
        const double va = A.pn * fm_exp_neg(-0.5 * za * za), vb = B.pn * fm_exp_neg(-0.5 * zb * zb);
        rowE[e] = live ? va : 0.;
        if (e + 1 < nE) rowE[e + 1] = live ? vb : 0.;
      }
      for (int n = 0; n < nN; n += 2) {
        const NClassDev A = s_n[n];
        const NClassDev B = s_n[min(n + 1, nN - 1)];
        // c0, c1 carry the 1/sqrt(2) of Phi(x) = 1 - erfc(x / sqrt(2)) / 2
        const double va = A.one ? 1. : fm_half_erfc(fma(A.c1, s, A.c0));
        const double vb = B.one ? 1. : fm_half_erfc(fma(B.c1, s, B.c0));
        rowN[n] = va;
        if (n + 1 < nN) rowN[n + 1] = vb;
      }
      __syncwarp();
      // phase 2: lane = product, dot product over the 32 event pairs
#pragma unroll
      for (int r = 0; r < CK_MAXR; ++r) {
        if (r < n_rounds) {
          const double* colE = myE + pe[r];
          const double* colN = myN + pn[r];
          // four independent chains (fixed order: deterministic)
          double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
#pragma unroll
          for (int l = 0; l < 32; l += 4) {
            a0 = fma(colE[l * sE], colN[l * sN], a0);
            a1 = fma(colE[(l + 1) * sE], colN[(l + 1) * sN], a1);
            a2 = fma(colE[(l + 2) * sE], colN[(l + 2) * sN], a2);
            a3 = fma(colE[(l + 3) * sE], colN[(l + 3) * sN], a3);
          }
          acc[r] += (a0 + a1) + (a2 + a3);
        }
      }
      __syncwarp();
      n_pairs += live ? 1 : 0;
    }
    fence_proxy_async();
    __syncthreads();
  }
  // ---- combine the warps' accumulators, one partial per work unit ---------------------------
  __syncthreads();
  double* s_red = s_tiles;
#pragma unroll
  for (int r = 0; r < CK_MAXR; ++r) s_red[(warp * CK_MAXR + r) * 32 + lane] = acc[r];
  __syncthreads();
  for (int q = tid; q < nProd; q += CK_NT) {
    const int r = q >> 5, l = q & 31;
    double v = 0.;
#pragma unroll
    for (int w = 0; w < CK_NT / 32; ++w) v += s_red[(w * CK_MAXR + r) * 32 + l];
    partials[T.part_first + (int64_t)chunk * nProd + q] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_pairs += __shfl_xor_sync(0xffffffffu, n_pairs, o);
  __shared__ unsigned long long s_cnt[CK_NT / 32];
  if (lane == 0) s_cnt[warp] = n_pairs;
  __syncthreads();
  if (tid == 0) {
    unsigned long long c = 0;
    for (int w = 0; w < CK_NT / 32; ++w) c += s_cnt[w];
    counts[wu] = c;
  }
}

// raw[dst] += scale * sum over the term's work units of partial[prod], in order.
// Every (term, product) owns its destination: no two threads write the same slot.
__global__ void class_finalize_kernel(const ClassTermDev* __restrict__ terms, int n_terms,
                                      const ProdDev* __restrict__ prods, const double* __restrict__ partials,
                                      double* __restrict__ raw) {
  const int t = blockIdx.x;
  if (t >= n_terms) return;
  const ClassTermDev T = terms[t];
  for (int q = threadIdx.x; q < T.nProd; q += blockDim.x) {
    double v = 0.;
    for (int w = 0; w < T.n_chunks; ++w) v += partials[T.part_first + (int64_t)w * T.nProd + q];
    const ProdDev P = prods[T.prod_first + q];
    raw[P.dst] += P.scale * v;
  }
}
