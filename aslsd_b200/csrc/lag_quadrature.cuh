// Lag quadrature: the interior of every exact double sum in one pass over the event pairs.
//
// For a target m far enough from T_f every double-sum term of the LSE,
//     sum_{m} sum_{n : t_n < t_m} f(t_m - t_n),           f = phi, d phi, Upsilon(+inf, .), d Upsilon(+inf, .)
// (mhp_exact_estimator.py:611-707 with the closed forms of basis_kernel_gaussian.py /
// basis_kernel_exponential.py), depends on the pair only through the lag s = t_m - t_n, and every f is
// an entire function of s whose scale (beta, sigma, 1/beta ...) is known.  All terms that share the
// ordered pair of event streams (target type, source type) therefore sum *different functions over the
// same multiset of lags*.  That multiset is compressed once into a quadrature rule:
//
//   1. lag_moment_kernel   [0, s_max] is cut into J bins of width h (a fraction of the smallest scale);
//                          with u in [-1, 1] the position of a lag inside its bin, the kernel accumulates
//                          the Chebyshev moments  M[b][d] = sum_{pairs in bin b} T_d(u),  d < LQ_ND.
//   2. lag_weights_kernel  W[b][i] = sum_d a_d T_d(x_i) M[b][d]  at the LQ_ND Chebyshev nodes x_i of the
//                          bin: sum_pairs p(s) = sum_{b,i} W[b][i] p(s_{b,i}) exactly for every piecewise
//                          polynomial p of degree < LQ_ND.
//   3. lag_quad_kernel     every row (one basis / basis pair of one term) evaluates the reference's
//                          closed form at the J * LQ_ND nodes and takes the dot product with W.
//
// Degree-11 interpolation on bins of at most 0.5 scale units reproduces these functions to ~1e-13 of
// their maximum (scripts/quadrature_error.py), far inside the 1e-10 parity bar, and steps 1-2 do not depend on
// the parameters at all.  The O(pairs) work drops from one exp + one erfc per (pair, basis pair) to ~2
// fp64 instructions per (pair, Chebyshev degree); the closed forms are evaluated O(rows * J * LQ_ND) times.
//
// lag_moment_kernel, one CTA = (stream, block of targets), per tile of NT targets (one per thread; NT, NW, GS: LqShape):
//   a. the tile's source window sits in shared memory, fetched by one TMA bulk copy that was issued
//      during the previous tile (the first source a tile can need is the window start of the previous
//      tile's last target: no global search, two buffers); per-target source ranges by binary search
//      in that window (bounded by the part of it the tile can reach; for the usual 511 / 1023
//      entries unrolled, without bound tests, the two searches of a target interleaved); exclusive
//      scan of the pair counts.  A window that does not fit falls back to bucket-table searches and
//      global reads for that tile;
//   b. counting sort of the tile's lags by bin, in chunks of `ucap` pairs.  Every target's thread
//      writes the lags of its pairs target-major into shared memory, so that afterwards every lane
//      of every warp holds a pair.  Pass 1 (two warp steps interleaved, full steps without lane
//      masks, the partial last step peeled): bin and in-bin coordinate
//      of each pair, rank among the equal bins of the warp step from ballots over the bits of the
//      bin index (match.any measured ~300 cycles: XU pipe; shared-memory atomics 2 cycles per
//      lane), running count in a warp-private histogram -> (bin, rank in warp) key.  Scan of the
//      histograms.  Pass 2: coordinates move to their sorted slots.  No atomics: the order inside a
//      bin is fixed and the result is deterministic;
//   c. every group of GS lanes takes a bin: three-term recurrence, LQ_ND - 1 register accumulators per lane,
//      transposed shuffle reduction, owner adds into the CTA's moment table.  With at most
//      NW * 32 / GS = 64 bins every group owns one bin for the whole work unit: its sums stay
//      in registers across chunks and tiles and are reduced once, before the partial table is written;
//   one partial table per work unit, reduced in fixed order by lag_weights_kernel.
#pragma once

#include "basis_math.cuh"
#include "pair_kernels.cuh"

#define LQ_UCAP_MAX 8192          // pairs sorted at a time
#define LQ_ND 12                  // Chebyshev terms = nodes per bin
#define LQ_JMAX 1024              // bins per stream
#define LQ_SCALE_FRACTION 0.5     // largest bin width in units of the smallest scale of the stream's rows
#define LQ_MAGIC 6755399441055744.0   // 1.5 * 2^52: x + MAGIC holds rint(x) in its low word

// Shape of the moment kernel's CTA: NT threads (one target of a tile each), CTAS of them resident per SM
// (the shared-memory budget is cut for that).  Two instances are compiled:
//   LqBig   512 x 1  serves any bin count up to LQ_JMAX (220 KB of shared memory);
//   LqSmall 128 x 4  for unweighted streams of at most 64 bins (every pair of lanes owns one bin): the four
//           CTAs of an SM are in different phases (sort: integer / shared-memory issue; recurrences:
//           fp64 pipe; barriers), which overlap - measured 10.45 ms against 12.6 ms per C5 step
//           (profiles/r2_variants.md; 256 x 2: 10.9 ms, 512 x 2 with 64 registers: 16.3 ms, 256 x 3: 12.3 ms).
template <int NT_, int CTAS_>
struct LqShape {
  static constexpr int NT = NT_, CTAS = CTAS_;
  static constexpr int NW = NT / 32;
  // lanes per bin in the moment pass (one round of the CTA covers 64 bins)
  static constexpr int GS = NT >= 1024 ? 16 : (NT >= 512 ? 8 : (NT >= 256 ? 4 : 2));
  static constexpr int TM = NT;     // targets per tile: one per thread
  static constexpr int SMAX = NT >= 512 ? 2048 : ((NT >= 256 && CTAS < 3) ? 1024 : (CTAS >= 5 ? 256 : 512));   // staged source events per tile (two buffers)
  // (227 KB per SM, 1 KB of each resident CTA reserved by the system)
  static constexpr int SMEM_BUDGET = CTAS == 1 ? 220 * 1024 : (CTAS == 2 ? 110 * 1024 : (CTAS == 3 ? 73 * 1024 : ((227 - CTAS) * 1024) / CTAS));
  static constexpr int ONE_BIN_J = NW * (32 / GS);   // bins the groups of lanes can own for a whole work unit
  // the small shape only takes streams whose bins are all owned (ONE_BIN_J): the moment table is touched
  // once, after the last tile, and shares its shared memory with the pair buffers
  static constexpr bool ALIAS_M = CTAS >= 4;
};
typedef LqShape<512, 1> LqBig;
#ifndef LQ_SMALL_CTAS
#define LQ_SMALL_CTAS 4
#endif
#ifndef LQ_SMALL_NT
#define LQ_SMALL_NT 128
#endif
typedef LqShape<LQ_SMALL_NT, LQ_SMALL_CTAS> LqSmall;
enum { LQ_SHAPE_BIG = 0, LQ_SHAPE_SMALL = 1, LQ_SHAPE_WARP = 2 };

struct QStreamDev {
  int32_t tgt_dim, src_dim, same, J;
  int32_t n_chunks, ucap;  // work units; pairs sorted at a time (multiple of the shape's NT)
  int32_t nbits, shape;    // bits of the bin index; LQ_SHAPE_* of the CTAs that take this stream
  int32_t tm, pad;         // targets per tile (<= the shape's NT): sized so that the pairs of a tile fit one chunk
  int64_t m_lo, m_hi;      // interior target block
  int64_t tpw;             // targets per work unit
  int64_t wu_first;        // first work unit (prefix over streams)
  int64_t part_first;      // partial moment tables: [chunk][J][LQ_ND] doubles
  int64_t w_first;         // weights: [J][LQ_ND] doubles
  int64_t wa_off, wb_off;  // per-event weights of targets / sources (marked models), -1 = none:
                           // the moments become sum of weight(m) weight(n) T_d(u)
  double s_max, h, inv_h;
};

// x_i = cos(pi (i + 1/2) / ND) and a_d T_d(x_i) (a_0 = 1/ND, a_d = 2/ND), filled by the host
__constant__ double LQ_X[LQ_ND];
__constant__ double LQ_CT[LQ_ND * LQ_ND];   // [i][d]

// what pass 1 needs of a tile's target to form the lags of its pairs: the pairs of a tile are numbered
// target-major, pair q of the chunk starting at pair c0 belongs to the target with pre <= c0 + q <
// pre + cnt and pairs it with window entry base + c0 + q
struct __align__(16) LqTarget {
  double t;
  int32_t base;   // (first source - window start) - pre
  int32_t pad;
};
#define LQ_HEAD 0x80000000u   // s_key entries with this bit: 1 + the target that owns the pair (pass 1 input)

// dynamic shared memory: [pair-order coordinates | sorted coordinates | 2 source windows |
//                         moment table | target records | keys | bin starts | histograms | cursors]
template <class C>
__host__ __device__ inline size_t lq_smem_bytes(int J, int ucap, bool weighted) {
  size_t b = 0;
  b += (size_t)ucap * sizeof(double) * (weighted ? 4 : 2);   // + pair weights in both orders
  b += (size_t)(C::SMAX + 2) * sizeof(double) * 2;
  if (!C::ALIAS_M) b += (size_t)J * LQ_ND * sizeof(double);
  b += (size_t)ucap * sizeof(uint32_t);
  b += (size_t)(J + 2) * sizeof(int32_t);
  b += (size_t)C::NW * J * sizeof(uint16_t) * (J <= 64 ? 2 : 1);
  b += (size_t)C::NT * (sizeof(LqTarget) + (weighted ? sizeof(double) : 0));
  return b + 64;
}

// largest chunk (multiple of C::NT) that fits the budget; 0 if none does
template <class C>
inline int lq_pick_ucap(int J, bool weighted) {
  for (int u = LQ_UCAP_MAX; u >= C::NT; u -= 32)
    if (lq_smem_bytes<C>(J, u, weighted) <= (size_t)C::SMEM_BUDGET) return u;
  return 0;
}

__device__ __forceinline__ int find_stream(const QStreamDev* __restrict__ streams, int n, int64_t wu) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (streams[mid].wu_first <= wu) lo = mid; else hi = mid;
  }
  return lo;
}

// first index in [0, n) of the sorted shared array with win[idx] >= x (n if none);
// p2 = largest power of two <= n (warp-uniform trip count, no data-dependent branch)
__device__ __forceinline__ int smem_lower_bound(const double* win, int n, int p2, double x) {
  int pos = 0;
  for (int step = p2; step > 0; step >>= 1) {
    const int np = pos + step;
    const double v = win[min(np, n) - 1];
    pos = (np <= n && v < x) ? np : pos;
  }
  return pos;
}

// the same over exactly 2^K - 1 entries: no bounds to test, fully unrolled
template <int K>
__device__ __forceinline__ int smem_lower_bound_full(const double* win, double x) {
  int pos = 0;
#pragma unroll
  for (int step = 1 << (K - 1); step > 0; step >>= 1)
    if (win[pos + step - 1] < x) pos += step;
  return pos;
}

// two such searches at once: the two dependent chains of shared-memory loads overlap
template <int K>
__device__ __forceinline__ void smem_lower_bound_full2(const double* win, double xa, double xb, int& pa, int& pb) {
  pa = 0;
  pb = 0;
#pragma unroll
  for (int step = 1 << (K - 1); step > 0; step >>= 1) {
    const double va = win[pa + step - 1], vb = win[pb + step - 1];
    if (va < xa) pa += step;
    if (vb < xb) pb += step;
  }
}

// N such searches at once
template <int K, int N>
__device__ __forceinline__ void smem_lower_bound_fullN(const double* win, const double* x, int* p) {
#pragma unroll
  for (int c = 0; c < N; ++c) p[c] = 0;
#pragma unroll
  for (int step = 1 << (K - 1); step > 0; step >>= 1) {
    double v[N];
#pragma unroll
    for (int c = 0; c < N; ++c) v[c] = win[p[c] + step - 1];
#pragma unroll
    for (int c = 0; c < N; ++c)
      if (v[c] < x[c]) p[c] += step;
  }
}

// lanes of the warp whose b equals this lane's, from ballots over NB bits
// (per bit: test, vote, select, and-not-xor)
template <int K>
__device__ __forceinline__ void match_bit(int b, unsigned& peers) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 t, v, m;\n"
      "and.b32 t, %1, %2;\n"
      "setp.ne.u32 p, t, 0;\n"
      "vote.sync.ballot.b32 v, p, 0xffffffff;\n"
      "selp.b32 m, 0xffffffff, 0, p;\n"
      "lop3.b32 %0, %0, v, m, 0x90;\n"   // peers & ~(v ^ m)
      "}\n"
      : "+r"(peers)
      : "r"(b), "n"(1 << K));
}
template <int NB>
__device__ __forceinline__ unsigned match_bins(int b) {
  unsigned peers = 0xffffffffu;
  match_bit<0>(b, peers); match_bit<1>(b, peers); match_bit<2>(b, peers); match_bit<3>(b, peers);
  match_bit<4>(b, peers);
  if (NB > 5) match_bit<5>(b, peers);
  if (NB > 6) match_bit<6>(b, peers);
  if (NB > 7) { match_bit<7>(b, peers); match_bit<8>(b, peers); match_bit<9>(b, peers); }
  return peers;
}

// bin index and position inside the bin (v in [-1/2, 1/2]; the Chebyshev variable is 2 v)
__device__ __forceinline__ void lq_bin(double s, double inv_h, int& b, double& v) {
  const double xb = fma(s, inv_h, -0.5);
  const double tr = xb + LQ_MAGIC;
  b = __double2loint(tr);
  v = xb - (tr - LQ_MAGIC);
}

// Pass 1 of the counting sort over the n_c pairs of a chunk (see the file header).  On entry s_key
// marks, for every target, the first of its pairs in the chunk and every 32nd slot after it
// (LQ_HEAD | 1 + target): a lane finds the owner of its pair from a ballot of the marks of its warp
// step (lane 0 always carries one), forms the lag from the target's record and the source window,
// and leaves (bin, rank in warp) in s_key and the position inside the bin in s_val.  No pass over the
// pairs by the targets' threads, whose pair counts differ (measured: 37 % of the instructions of the
// kernel before).  Two warp steps per iteration: the ballots of the two are independent, the
// histogram updates are ordered.
template <bool STAGED, bool WEIGHTED>
__device__ __forceinline__ double lq_pair_lag(uint32_t own, unsigned le_mask, int q, int c0,
                                              const LqTarget* __restrict__ s_tgt, const double* __restrict__ s_tw,
                                              const double* __restrict__ win, const double* __restrict__ src_glb,
                                              const double* __restrict__ w_src, double& w) {
  const unsigned heads = __ballot_sync(0xffffffffu, (int32_t)own < 0);
  const int hl = 31 - __clz(heads & le_mask);
  const int m = (int)(__shfl_sync(0xffffffffu, own, hl) & 0xffffu) - 1;
  const LqTarget T = s_tgt[m];
  const int idx = T.base + c0 + q;
  if (WEIGHTED) {
    w = s_tw[m];
    if (w_src) w *= w_src[idx];
  }
  return T.t - (STAGED ? win[idx] : __ldg(src_glb + idx));
}

template <int NB, bool STAGED, bool WEIGHTED, class C>
__device__ __forceinline__ void lq_pass1(int n_c, int c0, double inv_h, uint16_t* my_hist, uint32_t* s_key,
                                         double* s_val, double* s_wv, const LqTarget* __restrict__ s_tgt,
                                         const double* __restrict__ s_tw, const double* __restrict__ win,
                                         const double* __restrict__ src_glb, const double* __restrict__ w_src) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned le_mask = lt_mask | (1u << lane);
  const int n_steps = (n_c + 31) >> 5;
  const int n_full = n_c >> 5;   // steps whose 32 pairs all exist
  int st = warp;
  // full steps, two per iteration: no lane masks, no liveness tests
  for (; st + C::NW < n_full; st += 2 * C::NW) {
    const int qa = st * 32 + lane, qb = qa + C::NW * 32;
    const uint32_t oa = s_key[qa], ob = s_key[qb];
    double wa = 1., wb = 1.;
    const double sa = lq_pair_lag<STAGED, WEIGHTED>(oa, le_mask, qa, c0, s_tgt, s_tw, win, src_glb, w_src, wa);
    const double sb = lq_pair_lag<STAGED, WEIGHTED>(ob, le_mask, qb, c0, s_tgt, s_tw, win, src_glb, w_src, wb);
    int ba, bb;
    double va, vb;
    lq_bin(sa, inv_h, ba, va);
    lq_bin(sb, inv_h, bb, vb);
    const unsigned pa = match_bins<NB>(ba);
    const unsigned pb = match_bins<NB>(bb);
    const int ra = __popc(pa & lt_mask), rb = __popc(pb & lt_mask);
    const int base_a = my_hist[ba];
    __syncwarp();
    if (ra == 0) my_hist[ba] = (uint16_t)(base_a + __popc(pa));
    s_key[qa] = (uint32_t)ba | ((uint32_t)(base_a + ra) << 10);
    s_val[qa] = va;
    if (WEIGHTED) s_wv[qa] = wa;
    __syncwarp();
    const int base_b = my_hist[bb];
    __syncwarp();
    if (rb == 0) my_hist[bb] = (uint16_t)(base_b + __popc(pb));
    s_key[qb] = (uint32_t)bb | ((uint32_t)(base_b + rb) << 10);
    s_val[qb] = vb;
    if (WEIGHTED) s_wv[qb] = wb;
    __syncwarp();
  }
  // the remaining steps of this warp (at most two, the last one of the chunk possibly partial):
  // idle lanes redo the last pair and are masked out of every peer set
  for (; st < n_steps; st += C::NW) {
    const int qa = st * 32 + lane;
    const bool live_a = qa < n_c;
    const int left_a = n_c - st * 32;
    const unsigned mask_a = left_a >= 32 ? 0xffffffffu : ((1u << left_a) - 1u);
    const uint32_t oa = live_a ? s_key[qa] : 0u;
    double wa = 1.;
    const double sa = lq_pair_lag<STAGED, WEIGHTED>(oa, le_mask, live_a ? qa : n_c - 1, c0, s_tgt, s_tw, win, src_glb,
                                                    w_src, wa);
    int ba;
    double va;
    lq_bin(sa, inv_h, ba, va);
    const unsigned pa = match_bins<NB>(ba) & mask_a;
    const int ra = __popc(pa & lt_mask);
    const int base_a = my_hist[ba];
    __syncwarp();
    if (live_a) {
      if (ra == 0) my_hist[ba] = (uint16_t)(base_a + __popc(pa));
      s_key[qa] = (uint32_t)ba | ((uint32_t)(base_a + ra) << 10);
      s_val[qa] = va;
      if (WEIGHTED) s_wv[qa] = wa;
    }
    __syncwarp();
  }
}

// Chebyshev moments of the sorted positions s_u[e0 + qg], s_u[e0 + qg + C::GS], ... < e1, added to
// acc[1..] (acc[0], the count, is the caller's: exact for unweighted streams).  Two independent
// recurrences per iteration (the chain is latency-bound).
template <bool WEIGHTED, class C>
__device__ __forceinline__ void lq_accumulate(const double* __restrict__ s_u, const double* __restrict__ s_wu,
                                              int32_t e0, int32_t e1, int qg, double* acc) {
  int32_t e = e0 + qg;
  if (WEIGHTED) {
    for (; e < e1; e += C::GS) {
      const double w = s_wu[e];
      const double uu = s_u[e] + s_u[e];
      const double two_u = uu + uu;
      double tp = 1., tc = uu;
      acc[0] += w;
      acc[1] = fma(w, tc, acc[1]);
#pragma unroll
      for (int d = 2; d < LQ_ND; ++d) {
        const double tn = fma(two_u, tc, -tp);
        tp = tc;
        tc = tn;
        acc[d] = fma(w, tn, acc[d]);
      }
    }
  } else {
  for (; e + C::GS < e1; e += 2 * C::GS) {
    const double ua = s_u[e] + s_u[e], ub = s_u[e + C::GS] + s_u[e + C::GS];
    const double two_a = ua + ua, two_b = ub + ub;
    double pa = 1., ca = ua, pb = 1., cb = ub;
    acc[1] += ca + cb;
#pragma unroll
    for (int d = 2; d < LQ_ND; ++d) {
      const double na = fma(two_a, ca, -pa), nb = fma(two_b, cb, -pb);
      pa = ca; ca = na;
      pb = cb; cb = nb;
      acc[d] += na + nb;
    }
  }
  if (e < e1) {
    const double vv = s_u[e];
    const double uu = vv + vv;
    const double two_u = uu + uu;
    double tp = 1., tc = uu;
    acc[1] += tc;
#pragma unroll
    for (int d = 2; d < LQ_ND; ++d) {
      const double tn = fma(two_u, tc, -tp);
      tp = tc;
      tc = tn;
      acc[d] += tn;
    }
  }
  }
}

// Transposed reduction of acc[LQ_ND] over the C::GS lanes of a group (12 -> 6 -> 3 values per
// lane on the two widest exchanges, plain butterflies on the rest), then Mb[d] += the sums.
template <class C>
__device__ __forceinline__ void lq_reduce_store(double* acc, int qg, bool store, double* Mb) {
  {
    const bool up = (qg & (C::GS / 2)) != 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double send = up ? acc[i] : acc[i + 6];
      const double keep = up ? acc[i + 6] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, C::GS / 2);
    }
  }
  if constexpr (C::GS >= 4) {
  {
    const bool up = (qg & (C::GS / 4)) != 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double send = up ? acc[i] : acc[i + 3];
      const double keep = up ? acc[i + 3] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, C::GS / 4);
    }
  }
#pragma unroll
  for (int o = C::GS / 8; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  if (store && (qg & (C::GS / 4 - 1)) == 0) {
    const int d0 = ((qg & (C::GS / 2)) ? 6 : 0) + ((qg & (C::GS / 4)) ? 3 : 0);
    Mb[d0] += acc[0];
    Mb[d0 + 1] += acc[1];
    Mb[d0 + 2] += acc[2];
  }
  } else {
  // two lanes per bin: each ends up with six of the twelve sums
  if (store) {
    const int d0 = (qg & 1) ? 6 : 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) Mb[d0 + i] += acc[i];
  }
  }
}

// WEIGHTED: streams of marked models (per-event weights); launched separately from the
// unweighted streams so that those keep their register budget
template <bool WEIGHTED, class C>
__global__ void __launch_bounds__(C::NT, C::CTAS)
lag_moment_kernel(const DimDev* __restrict__ dims, const QStreamDev* __restrict__ streams, int n_streams,
                  double* __restrict__ partials, unsigned long long* __restrict__ counts,
                  const double* __restrict__ weights, int64_t wu_base) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ int64_t s_a0[2];        // first source index of each window buffer
  // barriers per tile: one after the pair counts of the warps, then per chunk one before pass 1, one
  // after it and one after pass 2 - what a tile writes before its first barrier is double-buffered
  __shared__ int64_t s_next_lo[2];   // window start of the tile's last target
  __shared__ int32_t s_warp[2][C::NW];   // pair counts of a tile fit 32 bits (host guard, build_stream)
  __shared__ int32_t s_scan[C::NW];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t wu = (int64_t)blockIdx.x + wu_base;
  const QStreamDev S = streams[find_stream(streams, n_streams, wu)];
  const int chunk = (int)(wu - S.wu_first);
  const DimDev Dt = dims[S.tgt_dim], Ds = dims[S.src_dim];
  if (*Dt.bad) return;   // unsorted upload (aslsd_set_events): the verdict reaches the host later
  const int64_t m_begin = S.m_lo + (int64_t)chunk * S.tpw;
  const int64_t m_end = min(S.m_hi, m_begin + S.tpw);
  const bool same = S.same != 0;
  const int J = S.J, ucap = S.ucap, nbits = S.nbits;
  const double inv_h = S.inv_h;
  const int64_t n_src_padded = ((Ds.n + 1) & ~(int64_t)1) + 2;   // set_events pads every type (ctx.cu)

  const bool weighted = WEIGHTED;
  double* s_val = (double*)smem_raw;
  double* s_u = s_val + ucap;
  double* s_wv = s_u + ucap;                        // pair weights, pair order (weighted streams)
  double* s_wu = s_wv + (weighted ? ucap : 0);      // pair weights, sorted by bin
  double* s_src = s_wu + (weighted ? ucap : 0);     // two buffers of C::SMAX + 2
  double* s_M = C::ALIAS_M ? s_val : s_src + 2 * (C::SMAX + 2);
  LqTarget* s_tgt = (LqTarget*)(s_src + 2 * (C::SMAX + 2) + (C::ALIAS_M ? 0 : (size_t)J * LQ_ND));   // 16-byte aligned: every block before it is
  double* s_tw = (double*)(s_tgt + C::NT);                  // target weights (weighted streams)
  uint32_t* s_key = (uint32_t*)(s_tw + (weighted ? C::NT : 0));
  int32_t* s_bin = (int32_t*)(s_key + ucap);
  uint16_t* s_hist = (uint16_t*)(s_bin + (J + 2));
  uint16_t* my_hist = s_hist + warp * J;
  uint16_t* s_cur = s_hist + C::NW * J;             // cursors of the warps' own scans (J <= 64)

  if (!C::ALIAS_M)
    for (int x = tid; x < J * LQ_ND; x += C::NT) s_M[x] = 0.;
  for (int x = tid; x < ucap; x += C::NT) s_key[x] = 0u;   // no marks (pass 1 leaves none behind)
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_proxy_async();
    // window of the first tile: starts at the first source its first target can reach
    int64_t a0 = 0;
    if (m_begin < m_end) a0 = dim_lower_bound(Ds, Dt.t[m_begin] - S.s_max) & ~(int64_t)1;
    s_a0[0] = a0;
    const int64_t n_stage = min((int64_t)C::SMAX, n_src_padded - a0);
    if (m_begin < m_end) {
      mbar_expect_tx(&s_bar[0], (uint32_t)(n_stage * sizeof(double)));
      tma_load_1d(s_src, Ds.t + a0, (uint32_t)(n_stage * sizeof(double)), &s_bar[0]);
    }
  }
  __syncthreads();
  uint32_t parity = 0u;   // bit b: phase of window buffer b
  unsigned long long n_pairs = 0;
  double t_mine = (m_begin + tid < m_end) ? Dt.t[m_begin + tid] : 0.;

  // bins per pass of the moment stage = groups of C::GS lanes in the CTA
  const int grp = lane / C::GS, qg = lane % C::GS;
  const bool one_bin = J <= C::ONE_BIN_J;
  const int my_bin = warp * (32 / C::GS) + grp;
  double racc[LQ_ND];
#pragma unroll
  for (int d = 0; d < LQ_ND; ++d) racc[d] = 0.;

  int tile = 0;
  const int tm = S.tm;
  for (int64_t m0 = m_begin; m0 < m_end; m0 += tm, ++tile) {
    const int nt = (int)min((int64_t)tm, m_end - m0);
    const int buf = tile & 1;
    double* win = s_src + buf * (C::SMAX + 2);
    // ---- a. per-target source ranges ---------------------------------------------------------
    const double t = t_mine;
    const double t_last = __ldg(Dt.t + m0 + nt - 1);
    if (m0 + tm + tid < m_end) t_mine = Dt.t[m0 + tm + tid];   // next tile's target, in flight
    // the window and its first index (written before the copy was issued) are visible to whoever
    // sees the copy complete
    mbar_wait(&s_bar[buf], (parity >> buf) & 1u);
    parity ^= 1u << buf;
    const int64_t a0 = s_a0[buf];
    const int n_win = (int)min((int64_t)C::SMAX, n_src_padded - a0);
    // the window serves the tile if it reaches the end of the stream or the last target (same type:
    // by index, ties in time included; other types: by time, the searches are lower bounds)
    const bool staged = (a0 + n_win >= Ds.n) ||
                        (same ? (a0 + n_win >= m0 + nt - 1) : (win[n_win - 1] >= t_last));
    const int n_valid = (int)min((int64_t)n_win, Ds.n - a0);   // entries that are events, not padding
    int64_t lo_abs = 0;
    int32_t cnt = 0;
    // every threshold of the tile is <= t_last: the searches stop at the first power of two (minus
    // one: one step less) whose entry reaches it
    int n_search = n_valid;
    if (staged) {
      if (C::TM <= 256 && n_valid > 255 && win[254] >= t_last) n_search = 255;
      else if (n_valid > 511 && win[510] >= t_last) n_search = 511;
      else if (n_valid > 1023 && win[1022] >= t_last) n_search = 1023;
    }
    const int p2 = n_search > 0 ? (1 << (31 - __clz(n_search))) : 0;
    if (tid < nt) {
      int64_t hi;
      if (staged) {
        const double x_lo = t - S.s_max;
        int r_hi, r_lo;
        if (n_search == 1023) {
          if (same) { r_hi = 0; r_lo = smem_lower_bound_full<10>(win, x_lo); }
          else smem_lower_bound_full2<10>(win, t, x_lo, r_hi, r_lo);
        } else if (C::TM <= 256 && n_search == 255) {
          if (same) { r_hi = 0; r_lo = smem_lower_bound_full<8>(win, x_lo); }
          else smem_lower_bound_full2<8>(win, t, x_lo, r_hi, r_lo);
        } else if (n_search == 511) {
          if (same) { r_hi = 0; r_lo = smem_lower_bound_full<9>(win, x_lo); }
          else smem_lower_bound_full2<9>(win, t, x_lo, r_hi, r_lo);
        } else {
          r_hi = same ? 0 : smem_lower_bound(win, n_search, p2, t);
          r_lo = smem_lower_bound(win, n_search, p2, x_lo);
        }
        hi = same ? (m0 + tid) : a0 + r_hi;
        lo_abs = min(a0 + (int64_t)r_lo, hi);
      } else {
        hi = same ? (m0 + tid) : dim_lower_bound(Ds, t);
        lo_abs = min(dim_lower_bound(Ds, t - S.s_max), hi);
      }
      cnt = (int32_t)(hi - lo_abs);
      if (tid == nt - 1) s_next_lo[buf] = lo_abs;
    }
    int32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[buf][warp] = incl;
    fence_proxy_async();   // this thread's generic reads of the other window (previous tile) before its refill
    __syncthreads();
    int32_t pre = incl - cnt, total = 0;
#pragma unroll
    for (int w = 0; w < C::NW; ++w) {
      const int32_t v = s_warp[buf][w];
      if (w < warp) pre += v;
      total += v;
    }
    // next tile's window into the other buffer (its readers, the previous tile's searches and pass 1,
    // are behind that tile's barriers)
    if (tid == 0 && m0 + tm < m_end) {
      const int64_t a1 = s_next_lo[buf] & ~(int64_t)1;
      s_a0[buf ^ 1] = a1;
      const int64_t n_stage = min((int64_t)C::SMAX, n_src_padded - a1);
      mbar_expect_tx(&s_bar[buf ^ 1], (uint32_t)(n_stage * sizeof(double)));
      tma_load_1d(s_src + (buf ^ 1) * (C::SMAX + 2), Ds.t + a1, (uint32_t)(n_stage * sizeof(double)),
                  &s_bar[buf ^ 1]);
    }
    n_pairs += (unsigned long long)cnt;
    // the target's record for pass 1 (the previous tile's readers are behind the barrier after its pass 1)
    {
      LqTarget rec;
      rec.t = t;
      rec.base = (int32_t)(lo_abs - a0) - pre;
      rec.pad = 0;
      s_tgt[tid] = rec;
      if (weighted) s_tw[tid] = (S.wa_off >= 0 && tid < nt) ? weights[S.wa_off + m0 + tid] : 1.;
    }
    const double* src_glb = Ds.t + a0;                                        // fallback tiles
    const double* w_src = (weighted && S.wb_off >= 0) ? weights + S.wb_off + a0 : nullptr;

    // ---- b / c. chunks of at most ucap pairs ------------------------------------------------------
    for (int32_t c0 = 0; c0 < total; c0 += ucap) {
      const int32_t n_c = min(ucap, total - c0);
      // marks for pass 1: the first pair of every target in the chunk and every 32nd slot it covers;
      // histograms to zero
      {
        const int32_t qa = max(pre - c0, 0), qb = min(pre + cnt - c0, n_c);
        if (qa < qb) {
          const uint32_t mark = LQ_HEAD | (uint32_t)(tid + 1);
          s_key[qa] = mark;
          for (int32_t q = (qa & ~31) + 32; q < qb; q += 32) s_key[q] = mark;
        }
      }
      for (int x = tid; x < (C::NW * J + 1) / 2; x += C::NT) ((uint32_t*)s_hist)[x] = 0u;
      __syncthreads();
      // pass 1: bin, position in the bin, rank among the equal bins seen so far by this warp
#define LQ_PASS1(NB, STAGED) \
  lq_pass1<NB, STAGED, WEIGHTED, C>(n_c, c0, inv_h, my_hist, s_key, s_val, s_wv, s_tgt, s_tw, win, src_glb, w_src)
      if (!staged) LQ_PASS1(10, false);
      else if (nbits <= 5) LQ_PASS1(5, true);
      else if (nbits == 6) LQ_PASS1(6, true);
      else if (nbits == 7) LQ_PASS1(7, true);
      else LQ_PASS1(10, true);
#undef LQ_PASS1
      __syncthreads();
      // bin totals -> bin starts; per-warp cursors = start + counts of the lower warps
      const uint16_t* my_cur = my_hist;
      if (J <= 64) {
        // few bins: every warp scans all of them for itself (lane: J / 32 adjacent bins) and keeps its
        // own cursors - no barrier between the scan and pass 2, which only moves the warp's own pairs
        const int bpl = J >> 5;   // J is a multiple of 32
        int32_t tot0 = 0, tot1 = 0, below0 = 0, below1 = 0;
#pragma unroll
        for (int w = 0; w < C::NW; ++w) {
          const uint32_t v = bpl == 2 ? ((const uint32_t*)(s_hist + w * J))[lane] : (uint32_t)s_hist[w * J + lane];
          const int32_t c_lo = (int32_t)(v & 0xffffu), c_hi = (int32_t)(v >> 16);
          if (w < warp) { below0 += c_lo; below1 += c_hi; }
          tot0 += c_lo;
          tot1 += c_hi;
        }
        int32_t inc = tot0 + tot1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int32_t v = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += v;
        }
        const int32_t start0 = inc - tot0 - tot1, start1 = start0 + tot0;
        uint16_t* cur = s_cur + warp * J;
        if (bpl == 2) ((uint32_t*)cur)[lane] = (uint32_t)(start0 + below0) | ((uint32_t)(start1 + below1) << 16);
        else cur[lane] = (uint16_t)(start0 + below0);
        if (warp == 0) {
          s_bin[lane * bpl] = start0;
          if (bpl == 2) s_bin[lane * 2 + 1] = start1;
          if (lane == 31) s_bin[J] = inc;
        }
        my_cur = cur;
        __syncwarp();
      } else {
        // (thread tid owns the jpt consecutive bins from tid * jpt)
        const int jpt = (J + C::NT - 1) / C::NT;
        const bool owns_bins = warp * 32 * jpt < J;   // warp-uniform: the other warps only meet the barrier
        int32_t tot = 0, inc = 0;
        if (owns_bins) {
          for (int q = 0; q < jpt; ++q) {
            const int b = tid * jpt + q;
            if (b < J) {
#pragma unroll
              for (int w = 0; w < C::NW; ++w) tot += s_hist[w * J + b];
            }
          }
          inc = tot;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
          }
          if (lane == 31) s_scan[warp] = inc;
        }
        __syncthreads();
        if (owns_bins) {
          int32_t run = inc - tot;
          for (int w = 0; w < warp; ++w) run += s_scan[w];
          for (int q = 0; q < jpt; ++q) {
            const int b = tid * jpt + q;
            if (b < J) {
              s_bin[b] = run;
#pragma unroll
              for (int w = 0; w < C::NW; ++w) {
                const int32_t v = s_hist[w * J + b];
                s_hist[w * J + b] = (uint16_t)run;
                run += v;
              }
              if (b == J - 1) s_bin[J] = run;
            }
          }
        }
        __syncthreads();
      }
      // pass 2: coordinates to their sorted slots (pair q was ranked by warp (q / 32) % C::NW = warp)
      {
        int32_t q = tid;
        // four pairs per iteration: the dependent loads (key -> cursor of its bin) of the four overlap
        for (; q + 3 * C::NT < n_c; q += 4 * C::NT) {
          uint32_t key[4];
          double val[4], wv[4];
          int pos[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            key[x] = s_key[q + x * C::NT];
            val[x] = s_val[q + x * C::NT];
            if (weighted) wv[x] = s_wv[q + x * C::NT];
          }
#pragma unroll
          for (int x = 0; x < 4; ++x) pos[x] = (int)my_cur[key[x] & 1023u] + (int)(key[x] >> 10);
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            s_u[pos[x]] = val[x];
            if (weighted) s_wu[pos[x]] = wv[x];
          }
        }
        for (; q < n_c; q += C::NT) {
          const uint32_t key = s_key[q];
          const int b = (int)(key & 1023u);
          const int pos = (int)my_cur[b] + (int)(key >> 10);
          s_u[pos] = s_val[q];
          if (weighted) s_wu[pos] = s_wv[q];
        }
      }
      __syncthreads();
      // ---- c. moments of every bin: C::GS lanes per bin -------------------------------------------
      if (one_bin) {
        // every group of lanes owns one bin for the whole work unit: the partial sums stay in
        // registers across chunks and tiles and are reduced once, at the end
        int32_t e0 = 0, e1 = 0;
        if (my_bin < J) { e0 = s_bin[my_bin]; e1 = s_bin[my_bin + 1]; }
        lq_accumulate<WEIGHTED, C>(s_u, s_wu, e0, e1, qg, racc);
        if (!WEIGHTED && qg == 0) racc[0] += (double)(e1 - e0);   // T_0 = 1: the bin count, exact
      } else {
        for (int b0 = warp * (32 / C::GS); b0 < J; b0 += C::NW * (32 / C::GS)) {
          const int b = b0 + grp;
          int32_t e0 = 0, e1 = 0;
          if (b < J) { e0 = s_bin[b]; e1 = s_bin[b + 1]; }
          double acc[LQ_ND];
#pragma unroll
          for (int d = 0; d < LQ_ND; ++d) acc[d] = 0.;
          lq_accumulate<WEIGHTED, C>(s_u, s_wu, e0, e1, qg, acc);
          if (!WEIGHTED && qg == 0) acc[0] = (double)(e1 - e0);
          lq_reduce_store<C>(acc, qg, b < J && e1 > e0, s_M + (size_t)min(b, J - 1) * LQ_ND);
        }
      }
      // no barrier here: what the moments read (sorted coordinates, bin starts) is next written behind
      // the barrier after the next pass 1, what is written before it (marks, histograms, target
      // records) was last read before the barrier after pass 2
    }
  }
  if (C::ALIAS_M) {
    // the pair buffers are free now: the moment table takes their place
    __syncthreads();
    for (int x = tid; x < J * LQ_ND; x += C::NT) s_M[x] = 0.;
    __syncthreads();
  }
  if (one_bin) lq_reduce_store<C>(racc, qg, my_bin < J, s_M + (size_t)min(my_bin, J - 1) * LQ_ND);
  __syncthreads();
  // ---- one partial moment table per work unit ---------------------------------------------------
  double* out = partials + S.part_first + (int64_t)chunk * J * LQ_ND;
  for (int x = tid; x < J * LQ_ND; x += C::NT) out[x] = s_M[x];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_pairs += __shfl_xor_sync(0xffffffffu, n_pairs, o);
  __shared__ unsigned long long s_cnt[C::NW];
  if (lane == 0) s_cnt[warp] = n_pairs;
  __syncthreads();
  if (tid == 0) {
    unsigned long long c = 0;
    for (int w = 0; w < C::NW; ++w) c += s_cnt[w];
    counts[wu] = c;
  }
}

// ---- Warp-autonomous shape ---------------------------------------------------------------------------
// One warp = one work unit, no CTA barrier anywhere: unweighted streams of 32 or 64 bins.  The kernel
// above sorts the pairs of a chunk by bin through shared memory (two passes, four CTA barriers per
// tile) so that groups of lanes can run the recurrences of one bin with register accumulators; a third
// of its time is the fixed latency of its short phases.  Here lane L *owns* bin L of the
// current pass for the whole work unit (11 register accumulators) and the pairs are routed to their
// owners through per-bin stacks in shared memory.  A stream of 64 bins is taken in two passes over
// the work unit's targets - lags in the first 32 bins, then in the last 32 - so that a pass
// enumerates only its own pairs (the search bounds of the two overlap by a thousandth of a bin and a
// pair is kept by the pass that its computed bin names: an exact partition).
//   a. tile = 64 targets, two per lane: source window by a TMA bulk copy issued one tile ahead (two
//      buffers, one mbarrier each), per-target ranges by unrolled shared-memory searches, exclusive
//      scan by shuffles, one 16-byte record per target, head marks (one byte per pair slot);
//   b. step = 32 consecutive pairs (target-major), one per lane, two steps interleaved: owner of the
//      pair from a ballot over the marks, lag from the record and the window, bin and position in
//      the bin.  Five ballots over the bits of the bin (and one over "this pass") give every owner
//      lane the set of lanes that hold pairs of its bin (five LOP3 on the warp-uniform ballots); a
//      pair's lane fetches that set and the owner's stack height with two shuffles - rank among its
//      peers, slot - and pushes the position onto the bin's stack (s_stack[slot][bin]: the owners'
//      reads are conflict-free);
//   c. when a push would overflow a stack (LQW_CAP entries per bin) every owner pops the top I
//      entries of its stack - two independent three-term recurrences per lane, lanes whose stack is
//      shorter run with weight zero - with I chosen so that the fullest stack keeps LQW_RES free
//      slots.  Popping a fixed number from every stack instead of emptying them keeps ~88 % of the
//      lanes busy (emptying: 55 %; the sort above: 66 %); the summation order inside a bin is fixed,
//      no atomics.  A step that holds more pairs of one bin than a stack takes (ties, bursts) is fed
//      to the owners pair by pair.
// One partial moment table per work unit, written from the owners' registers.
#ifndef LQW_CAP
#define LQW_CAP 24      // stack entries per bin
#endif
#ifndef LQW_RES
#define LQW_RES 10      // free slots of the fullest stack after a partial drain
#endif
#ifndef LQW_PC
#define LQW_PC 1024     // pair slots with head marks (pairs of a tile are taken LQW_PC at a time)
#endif
#ifndef LQW_WARPS
#define LQW_WARPS 20    // resident warps (CTAs of one warp) per SM
#endif
#ifndef LQW_DEFAULT
#define LQW_DEFAULT 1   // 0: the 128-thread CTAs unless ASLSD_LQ_WARP=1
#endif
#ifndef LQW_NTL
#define LQW_NTL 2       // targets per lane and tile
#endif
#define LQW_TM (32 * LQW_NTL)   // targets per tile
#ifndef LQW_SW
#define LQW_SW (LQW_NTL >= 4 ? 256 : 128)   // staged source events per tile (two buffers)
#endif

// a value the compiler may not re-derive inside the loops (it otherwise rebuilds lane masks and
// shared-memory addresses from the lane index and the CTA's shared window in every step)
__device__ __forceinline__ void lqw_keep(unsigned& x) { asm volatile("" : "+r"(x)); }

// shared-memory accesses of the step loop by 32-bit address (the addresses stay in registers; the
// compiler otherwise re-derives every array base from the CTA's shared window in each iteration)
__device__ __forceinline__ uint32_t lqw_lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void lqw_sts_u8(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ double lqw_lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void lqw_sts_f64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void lqw_lds_v2(uint32_t a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}

// pop the top n_pop entries of this lane's stack (fewer if it is shorter): two recurrences at a time
__device__ __forceinline__ void lqw_drain(uint32_t col_a, int n_pop, int& cnt, double* acc, double& tot) {
  const int pop = min(cnt, n_pop);
  for (int i = 0; i < n_pop; i += 2) {
    const int ea = cnt - 1 - i, eb = cnt - 2 - i;
    const double wa = i < pop ? 1. : 0., wb = i + 1 < pop ? 1. : 0.;
    const double va = lqw_lds_f64(col_a + (uint32_t)max(ea, 0) * 256u), vb = lqw_lds_f64(col_a + (uint32_t)max(eb, 0) * 256u);
    const double ua = va + va, ub = vb + vb;
    const double two_a = ua + ua, two_b = ub + ub;
    double pa = 1., ca = ua, pb = 1., cb = ub;
    acc[1] = fma(wa, ca, acc[1]);
    acc[1] = fma(wb, cb, acc[1]);
#pragma unroll
    for (int d = 2; d < LQ_ND; ++d) {
      const double na = fma(two_a, ca, -pa), nb = fma(two_b, cb, -pb);
      pa = ca; ca = na;
      pb = cb; cb = nb;
      acc[d] = fma(wa, na, acc[d]);
      acc[d] = fma(wb, nb, acc[d]);
    }
  }
  tot += (double)pop;
  cnt -= pop;
}

// the pairs of a step straight into the accumulators of their owners (steps that no stack can take).
// Rare: a real function (the step loop stays small), the accumulators travel through local memory.
__device__ __noinline__ void lqw_direct_call(int lane, int b, double v, unsigned live, double* buf) {
  double acc[LQ_ND];
#pragma unroll
  for (int d = 1; d < LQ_ND; ++d) acc[d] = buf[d];
  double tot = buf[0];
  for (int k = 0; k < 32; ++k) {
    const int bk = __shfl_sync(0xffffffffu, b, k);
    const double vk = __shfl_sync(0xffffffffu, v, k);
    if (!((live >> k) & 1u)) continue;
    const double w = (bk & 31) == lane ? 1. : 0.;
    const double uu = vk + vk, two_u = uu + uu;
    double tp = 1., tc = uu;
    tot += w;
    acc[1] = fma(w, tc, acc[1]);
#pragma unroll
    for (int d = 2; d < LQ_ND; ++d) {
      const double tn = fma(two_u, tc, -tp);
      tp = tc;
      tc = tn;
      acc[d] = fma(w, tn, acc[d]);
    }
  }
#pragma unroll
  for (int d = 1; d < LQ_ND; ++d) buf[d] = acc[d];
  buf[0] = tot;
}
__device__ __forceinline__ void lqw_direct(int lane, int b, double v, unsigned live, double* acc, double& tot) {
  double buf[LQ_ND];
  buf[0] = tot;
#pragma unroll
  for (int d = 1; d < LQ_ND; ++d) buf[d] = acc[d];
  lqw_direct_call(lane, b, v, live, buf);
  tot = buf[0];
#pragma unroll
  for (int d = 1; d < LQ_ND; ++d) acc[d] = buf[d];
}

// ballot of bit K of b (one LOP3 with a predicate result + one VOTE)
template <int K>
__device__ __forceinline__ unsigned lqw_bit_ballot(int b) {
  unsigned v;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 t;\n"
      "and.b32 t, %1, %2;\n"
      "setp.ne.u32 p, t, 0;\n"
      "vote.sync.ballot.b32 %0, p, 0xffffffff;\n"
      "}\n"
      : "=r"(v)
      : "r"(b), "n"(1 << K));
  return v;
}

// what the step loop keeps in registers
struct LqwLoop {
  uint32_t mark_a;    // s_mark
  uint32_t tgt_a;     // s_tgt - 16 (marks hold 1 + the target)
  uint32_t win_a;     // the tile's source window
  uint32_t push_a;    // s_stack - 8 * 32 * pass: a live pair's bin b lies in [32 pass, 32 pass + 32)
  uint32_t col_a;     // s_stack + 8 * lane: this lane's stack
  const double* src_glb;   // fallback tiles: the sources from the window start on
  double inv_h;
  unsigned lt_mask, le_mask;
  unsigned xm[5];     // all ones where bit i of the lane index is 0
  int pass, lane;
};

// One pair per lane: lag -> (bin, position), whether this pass owns the bin, and for the owner lanes
// the set of lanes whose pair falls into their bin.  q: pair slot of this lane in the chunk; cq: its
// pair index in the tile.
template <bool FULL, bool STAGED>
__device__ __forceinline__ void lqw_form(const LqwLoop& L, int q, int cq, bool in_range, int& b, double& v,
                                         bool& live, unsigned& lv, unsigned& common) {
  uint32_t mk = 0u;
  if (FULL || in_range) mk = lqw_lds_u8(L.mark_a + (uint32_t)q);
  const unsigned heads = __ballot_sync(0xffffffffu, mk != 0u);
  const int hl = 31 - __clz(heads & L.le_mask);
  const int m1 = (int)__shfl_sync(0xffffffffu, mk, hl);   // 1 + owner (lane 0 of a step always carries a mark)
  if (FULL || in_range) lqw_sts_u8(L.mark_a + (uint32_t)q, 0u);
  double t_m, rest;
  lqw_lds_v2(L.tgt_a + ((uint32_t)m1 << 4), t_m, rest);
  int idx = __double2loint(rest) + cq;
  if (!FULL && !in_range) idx = 0;
  const double s = t_m - (STAGED ? lqw_lds_f64(L.win_a + ((uint32_t)idx << 3)) : __ldg(L.src_glb + idx));
  lq_bin(s, L.inv_h, b, v);
  live = ((b >> 5) == L.pass) && (FULL || in_range);
  lv = __ballot_sync(0xffffffffu, live);
  const unsigned B0 = lqw_bit_ballot<0>(b), B1 = lqw_bit_ballot<1>(b), B2 = lqw_bit_ballot<2>(b),
                 B3 = lqw_bit_ballot<3>(b), B4 = lqw_bit_ballot<4>(b);
  common = lv & (B0 ^ L.xm[0]) & (B1 ^ L.xm[1]) & (B2 ^ L.xm[2]) & (B3 ^ L.xm[3]) & (B4 ^ L.xm[4]);
}

// push a formed pair: rank among the lanes of the same bin, slot above the owner's height `cnt_at`
__device__ __forceinline__ void lqw_push(const LqwLoop& L, int b, double v, bool live, unsigned common, int cnt_at) {
  // (shuffles take the source lane modulo 32: the owner of bin b is lane b & 31)
  const unsigned pc = __shfl_sync(0xffffffffu, common, b);
  const int ck = __shfl_sync(0xffffffffu, cnt_at, b);
  const int slot = ck + __popc(pc & L.lt_mask);
  if (live) lqw_sts_f64(L.push_a + ((uint32_t)(slot * 32 + b) << 3), v);
}

// make room for n more entries on every stack (n: this lane's bin), or report that no drain can
template <class ACC>
__device__ __forceinline__ bool lqw_make_room(const LqwLoop& L, int n, int& cnt, ACC* acc, double& tot) {
  int need = cnt + n - LQW_CAP;
  need = __reduce_max_sync(0xffffffffu, need);
  const int top = __reduce_max_sync(0xffffffffu, cnt);
  // (an even number: the drain pops two entries per iteration)
  const int n_pop = min(top, (max(top - (LQW_CAP - LQW_RES), need) + 1) & ~1);
  __syncwarp();   // the pushes of the other lanes onto this lane's stack
  lqw_drain(L.col_a, n_pop, cnt, acc, tot);
  return !__any_sync(0xffffffffu, cnt + n > LQW_CAP);
}

// Two formed steps onto the stacks (the second one above the first).
__device__ __forceinline__ void lqw_push_two(const LqwLoop& L, int ba, double va, bool la, unsigned lva, unsigned ca,
                                             int bb, double vb, bool lb, unsigned lvb, unsigned cb, int& cnt,
                                             double* acc, double& tot) {
  const int na = __popc(ca), nb = __popc(cb);
  if (__any_sync(0xffffffffu, cnt + na + nb > LQW_CAP)) {
    if (!lqw_make_room(L, na + nb, cnt, acc, tot)) {
      // one step at a time, then pair by pair
      if (__any_sync(0xffffffffu, cnt + na > LQW_CAP) && !lqw_make_room(L, na, cnt, acc, tot))
        lqw_direct(L.lane, ba, va, lva, acc, tot);
      else {
        lqw_push(L, ba, va, la, ca, cnt);
        cnt += na;
      }
      if (__any_sync(0xffffffffu, cnt + nb > LQW_CAP) && !lqw_make_room(L, nb, cnt, acc, tot))
        lqw_direct(L.lane, bb, vb, lvb, acc, tot);
      else {
        lqw_push(L, bb, vb, lb, cb, cnt);
        cnt += nb;
      }
      return;
    }
  }
  lqw_push(L, ba, va, la, ca, cnt);
  lqw_push(L, bb, vb, lb, cb, cnt + na);
  cnt += na + nb;
}

// The steps over pair slots 0 .. n_c of a chunk (pairs c0 .. c0 + n_c of the tile), two at a time.
template <bool STAGED>
__device__ __forceinline__ void lqw_steps(const LqwLoop& L, int c0, int n_c, int& cnt, double* acc, double& tot) {
  int q = L.lane;
  int left = n_c;
  // full steps, two interleaved
  for (; left >= 64; left -= 64, q += 64) {
    int ba, bb;
    double va, vb;
    bool la, lb;
    unsigned lva, lvb, ca, cb;
    lqw_form<true, STAGED>(L, q, c0 + q, true, ba, va, la, lva, ca);
    lqw_form<true, STAGED>(L, q + 32, c0 + q + 32, true, bb, vb, lb, lvb, cb);
    lqw_push_two(L, ba, va, la, lva, ca, bb, vb, lb, lvb, cb, cnt, acc, tot);
  }
  if (left > 0) {
    // the chunk's last one or two steps (one copy of the code for every remainder)
    int ba, bb = 0;
    double va, vb = 0.;
    bool la, lb = false;
    unsigned lva, lvb = 0u, ca, cb = 0u;
    lqw_form<false, STAGED>(L, q, c0 + q, q < n_c, ba, va, la, lva, ca);
    if (left > 32) lqw_form<false, STAGED>(L, q + 32, c0 + q + 32, q + 32 < n_c, bb, vb, lb, lvb, cb);
    lqw_push_two(L, ba, va, la, lva, ca, bb, vb, lb, lvb, cb, cnt, acc, tot);
  }
}

// Tiles whose source window does not fit the staged buffer read their sources from global memory.
// Rare (a source stream far denser than the target stream): a real function, so that the kernel holds
// one copy of the step loop's code; loop state and accumulators travel through local memory.
__device__ __noinline__ void lqw_steps_global_call(const LqwLoop* Lp, int c0, int n_c, int* cnt_p, double* buf) {
  const LqwLoop L = *Lp;
  double acc[LQ_ND];
#pragma unroll
  for (int d = 1; d < LQ_ND; ++d) acc[d] = buf[d];
  double tot = buf[0];
  int cnt = *cnt_p;
  lqw_steps<false>(L, c0, n_c, cnt, acc, tot);
#pragma unroll
  for (int d = 1; d < LQ_ND; ++d) buf[d] = acc[d];
  buf[0] = tot;
  *cnt_p = cnt;
}
__device__ __forceinline__ void lqw_steps_global(const LqwLoop& L, int c0, int n_c, int& cnt, double* acc, double& tot) {
  LqwLoop Lc = L;
  double buf[LQ_ND];
  int c = cnt;
  buf[0] = tot;
#pragma unroll
  for (int d = 1; d < LQ_ND; ++d) buf[d] = acc[d];
  lqw_steps_global_call(&Lc, c0, n_c, &c, buf);
  cnt = c;
  tot = buf[0];
#pragma unroll
  for (int d = 1; d < LQ_ND; ++d) acc[d] = buf[d];
}

__global__ void __launch_bounds__(32, LQW_WARPS)
lag_moment_warp_kernel(const DimDev* __restrict__ dims, const QStreamDev* __restrict__ streams, int n_streams,
                       double* __restrict__ partials, unsigned long long* __restrict__ counts, int64_t wu_base) {
  __shared__ __align__(16) double s_stack[LQW_CAP * 32];
  __shared__ __align__(16) double s_src[2 * (LQW_SW + 2)];
  __shared__ __align__(16) LqTarget s_tgt[LQW_TM];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ __align__(16) uint8_t s_mark[LQW_PC];

  const int lane = threadIdx.x;
  const int64_t wu = (int64_t)blockIdx.x + wu_base;
  const QStreamDev S = streams[find_stream(streams, n_streams, wu)];
  const int chunk = (int)(wu - S.wu_first);
  const DimDev Dt = dims[S.tgt_dim], Ds = dims[S.src_dim];
  if (*Dt.bad) return;   // unsorted upload (aslsd_set_events): the verdict reaches the host later
  const int64_t m_begin = S.m_lo + (int64_t)chunk * S.tpw;
  const int64_t m_end = min(S.m_hi, m_begin + S.tpw);
  const int64_t n_src_padded = ((Ds.n + 1) & ~(int64_t)1) + 2;   // set_events pads every type (ctx.cu)
  const int n_pass = S.J >> 5;

  LqwLoop L;
  L.lane = lane;
  L.lt_mask = (1u << lane) - 1u;
  L.le_mask = L.lt_mask | (1u << lane);
#pragma unroll
  for (int i = 0; i < 5; ++i) L.xm[i] = ((lane >> i) & 1) ? 0u : 0xffffffffu;
  L.inv_h = S.inv_h;
  L.mark_a = smem_u32(s_mark);
  L.tgt_a = smem_u32(s_tgt) - 16u;
  L.col_a = smem_u32(s_stack) + 8u * (uint32_t)lane;
  lqw_keep(L.lt_mask);
  lqw_keep(L.le_mask);
#pragma unroll
  for (int i = 0; i < 5; ++i) lqw_keep(L.xm[i]);
  lqw_keep(L.mark_a);
  lqw_keep(L.tgt_a);
  lqw_keep(L.col_a);

  // stacks to zero (entries above a stack's height are read with weight zero: they must be finite),
  // no marks (every step clears the ones it read)
  for (int x = lane; x < LQW_CAP * 32; x += 32) s_stack[x] = 0.;
  for (int x = lane; x < LQW_PC / 4; x += 32) ((uint32_t*)s_mark)[x] = 0u;
  if (lane == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_proxy_async();
  }
  __syncwarp();
  uint32_t parity = 0u;   // bit b: phase of window buffer b
  unsigned long long n_pairs = 0;
  double* out = partials + S.part_first + (int64_t)chunk * S.J * LQ_ND;
  int tile = 0;

  for (int pass = 0; pass < n_pass; ++pass) {
    // lags of this pass: bins [32 pass, 32 pass + 32).  Sources of a target at t: t - s_hi <= t_n < t - s_lo
    // (first pass: t_n < t, by index for a stream paired with itself; last pass: down to t - s_max)
    const bool by_index = S.same != 0 && pass == 0;
    const double s_lo = pass == 0 ? 0. : (32. * pass - 1e-3) * S.h;
    const double s_hi = pass == n_pass - 1 ? S.s_max : (32. * (pass + 1) + 1e-3) * S.h;
    L.pass = pass;
    L.push_a = smem_u32(s_stack) - 256u * (uint32_t)pass;
    lqw_keep(L.push_a);
    double acc[LQ_ND];   // [0] unused: the count is kept as tot (exact)
#pragma unroll
    for (int d = 0; d < LQ_ND; ++d) acc[d] = 0.;
    double tot = 0.;
    int cnt = 0;   // height of this lane's stack

    int64_t a0 = 0;
    if (lane == 0 && m_begin < m_end) {
      // window of the first tile: starts at the first source its first target can reach
      a0 = dim_lower_bound(Ds, Dt.t[m_begin] - s_hi) & ~(int64_t)1;
      const int64_t n_stage = min((int64_t)LQW_SW, n_src_padded - a0);
      mbar_expect_tx(&s_bar[tile & 1], (uint32_t)(n_stage * sizeof(double)));
      tma_load_1d(s_src + (tile & 1) * (LQW_SW + 2), Ds.t + a0, (uint32_t)(n_stage * sizeof(double)), &s_bar[tile & 1]);
    }
    a0 = __shfl_sync(0xffffffffu, a0, 0);
    // this lane's targets of a tile: LQW_NTL consecutive ones
    double t_mine[LQW_NTL];
#pragma unroll
    for (int k = 0; k < LQW_NTL; ++k)
      t_mine[k] = (m_begin + LQW_NTL * lane + k < m_end) ? Dt.t[m_begin + LQW_NTL * lane + k] : 0.;

    for (int64_t m0 = m_begin; m0 < m_end; m0 += LQW_TM, ++tile) {
      const int nt = (int)min((int64_t)LQW_TM, m_end - m0);
      const int buf = tile & 1;
      const double* win = s_src + buf * (LQW_SW + 2);
      // ---- a. per-target source ranges ---------------------------------------------------------
      double t[LQW_NTL];
#pragma unroll
      for (int k = 0; k < LQW_NTL; ++k) t[k] = t_mine[k];
      // the tile's last target: lane (nt - 1) / LQW_NTL, its target (nt - 1) % LQW_NTL
      const int l_last = (nt - 1) / LQW_NTL, k_last = (nt - 1) % LQW_NTL;
      double t_last = __shfl_sync(0xffffffffu, t[0], l_last);
#pragma unroll
      for (int k = 1; k < LQW_NTL; ++k) {
        const double tk = __shfl_sync(0xffffffffu, t[k], l_last);
        if (k_last == k) t_last = tk;
      }
      const double x_last = t_last - s_lo;   // no threshold of the tile lies above
#pragma unroll
      for (int k = 0; k < LQW_NTL; ++k)   // next tile's targets, in flight
        if (m0 + LQW_TM + LQW_NTL * lane + k < m_end) t_mine[k] = Dt.t[m0 + LQW_TM + LQW_NTL * lane + k];
      mbar_wait(&s_bar[buf], (parity >> buf) & 1u);
      parity ^= 1u << buf;
      const int n_win = (int)min((int64_t)LQW_SW, n_src_padded - a0);
      // the window serves the tile if it reaches the end of the stream or the last target's sources
      // (by index: ties in time included; by time: the searches are lower bounds)
      const bool staged = (a0 + n_win >= Ds.n) ||
                          (by_index ? (a0 + n_win >= m0 + nt - 1) : (win[n_win - 1] >= x_last));
      const int n_valid = (int)min((int64_t)n_win, Ds.n - a0);   // entries that are events, not padding
      int n_search = n_valid;
      if (staged) {
        if (n_valid > 63 && win[62] >= x_last) n_search = 63;
        else if (n_valid > 127 && win[126] >= x_last) n_search = 127;
        else if (LQW_SW > 128 && n_valid > 255 && win[254] >= x_last) n_search = 255;
      }
      const int p2 = n_search > 0 ? (1 << (31 - __clz(n_search))) : 0;
      int64_t lo_abs[LQW_NTL];
      int32_t npair[LQW_NTL];
      {
        // thresholds: [0, NTL) lower ends, [NTL, 2 NTL) upper ends (not needed by index)
        double x[2 * LQW_NTL];
        int r[2 * LQW_NTL];
#pragma unroll
        for (int k = 0; k < LQW_NTL; ++k) {
          x[k] = t[k] - s_hi;
          x[LQW_NTL + k] = t[k] - s_lo;
          r[k] = r[LQW_NTL + k] = 0;
        }
        if (staged) {
          if (n_search == 63) {
            if (by_index) smem_lower_bound_fullN<6, LQW_NTL>(win, x, r);
            else smem_lower_bound_fullN<6, 2 * LQW_NTL>(win, x, r);
          } else if (n_search == 127) {
            if (by_index) smem_lower_bound_fullN<7, LQW_NTL>(win, x, r);
            else smem_lower_bound_fullN<7, 2 * LQW_NTL>(win, x, r);
          } else if (LQW_SW > 128 && n_search == 255) {
            if (by_index) smem_lower_bound_fullN<8, LQW_NTL>(win, x, r);
            else smem_lower_bound_fullN<8, 2 * LQW_NTL>(win, x, r);
          } else {
#pragma unroll
            for (int k = 0; k < LQW_NTL; ++k) {
              r[k] = smem_lower_bound(win, n_search, p2, x[k]);
              if (!by_index) r[LQW_NTL + k] = smem_lower_bound(win, n_search, p2, x[LQW_NTL + k]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < LQW_NTL; ++k) {
          const int j = LQW_NTL * lane + k;   // target of the tile
          int64_t hi, lo;
          if (staged) {
            hi = by_index ? (m0 + j) : a0 + r[LQW_NTL + k];
            lo = a0 + (int64_t)r[k];
          } else if (j < nt) {
            hi = by_index ? (m0 + j) : dim_lower_bound(Ds, x[LQW_NTL + k]);
            lo = dim_lower_bound(Ds, x[k]);
          } else {
            hi = lo = a0;
          }
          lo = min(lo, hi);
          lo_abs[k] = lo;
          npair[k] = j < nt ? (int32_t)(hi - lo) : 0;
        }
      }
      int32_t mine = 0;
#pragma unroll
      for (int k = 0; k < LQW_NTL; ++k) mine += npair[k];
      int32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int32_t total = __shfl_sync(0xffffffffu, incl, 31);
      // next tile's window into the other buffer: it starts at the window start of this tile's last
      // target (the readers of that buffer, the previous tile's searches and steps, are done)
      int64_t a1 = __shfl_sync(0xffffffffu, lo_abs[0], l_last);
#pragma unroll
      for (int k = 1; k < LQW_NTL; ++k) {
        const int64_t ak = __shfl_sync(0xffffffffu, lo_abs[k], l_last);
        if (k_last == k) a1 = ak;
      }
      a1 &= ~(int64_t)1;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && m0 + LQW_TM < m_end) {
        const int64_t n_stage = min((int64_t)LQW_SW, n_src_padded - a1);
        mbar_expect_tx(&s_bar[buf ^ 1], (uint32_t)(n_stage * sizeof(double)));
        tma_load_1d(s_src + (buf ^ 1) * (LQW_SW + 2), Ds.t + a1, (uint32_t)(n_stage * sizeof(double)), &s_bar[buf ^ 1]);
      }
      int32_t pre[LQW_NTL];
      {
        int32_t run = incl - mine;
#pragma unroll
        for (int k = 0; k < LQW_NTL; ++k) {
          pre[k] = run;
          run += npair[k];
          LqTarget rec;
          rec.t = t[k];
          rec.base = (int32_t)(lo_abs[k] - a0) - pre[k];
          rec.pad = 0;
          s_tgt[LQW_NTL * lane + k] = rec;
        }
      }
      L.win_a = smem_u32(win);
      lqw_keep(L.win_a);
      L.src_glb = Ds.t + a0;   // fallback tiles

      // ---- b. steps of 32 pairs, LQW_PC pair slots at a time --------------------------------------
      for (int32_t c0 = 0; c0 < total; c0 += LQW_PC) {
        const int32_t n_c = min((int32_t)LQW_PC, total - c0);
#pragma unroll
        for (int k = 0; k < LQW_NTL; ++k) {
          const int32_t qa = max(pre[k] - c0, 0), qb = min(pre[k] + npair[k] - c0, n_c);
          if (qa < qb) {
            const uint8_t mark = (uint8_t)(LQW_NTL * lane + k + 1);
            s_mark[qa] = mark;
            // ... and every 32nd slot after it (one at most unless a target has more than 32 pairs)
            const int32_t q1 = (qa & ~31) + 32;
            if (q1 < qb) {
              s_mark[q1] = mark;
              if (q1 + 32 < qb)
                for (int32_t q = q1 + 32; q < qb; q += 32) s_mark[q] = mark;
            }
          }
        }
        __syncwarp();
        if (staged) lqw_steps<true>(L, c0, n_c, cnt, acc, tot);
        else lqw_steps_global(L, c0, n_c, cnt, acc, tot);
        __syncwarp();
      }
      a0 = a1;
    }
    // what is left on the stacks
    {
      const int top = __reduce_max_sync(0xffffffffu, cnt);
      __syncwarp();
      lqw_drain(L.col_a, top, cnt, acc, tot);
    }
    // ---- this pass's 32 bins of the work unit's partial moment table ---------------------------
    double* o = out + (size_t)(32 * pass + lane) * LQ_ND;
    o[0] = tot;
#pragma unroll
    for (int d = 1; d < LQ_ND; ++d) o[d] = acc[d];
    n_pairs += (unsigned long long)tot;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_pairs += __shfl_xor_sync(0xffffffffu, n_pairs, o);
  if (lane == 0) counts[wu] = n_pairs;
}

// Moments of a stream (partials summed in work-unit order) -> node weights.
// grid = (bins in groups of LQ_WB, streams); block = LQ_WB * LQ_ND threads.
#define LQ_WB 16
__global__ void __launch_bounds__(LQ_WB * LQ_ND)
lag_weights_kernel(const QStreamDev* __restrict__ streams, const double* __restrict__ partials,
                   double* __restrict__ weights) {
  __shared__ double s_m[LQ_WB][LQ_ND];
  const QStreamDev S = streams[blockIdx.y];
  const int jb = threadIdx.x / LQ_ND, c = threadIdx.x % LQ_ND;
  const int j = blockIdx.x * LQ_WB + jb;
  if (blockIdx.x * LQ_WB >= S.J) return;
  double v = 0.;
  if (j < S.J) {
    const double* p = partials + S.part_first + (int64_t)j * LQ_ND + c;
    const int64_t stride = (int64_t)S.J * LQ_ND;
    for (int w = 0; w < S.n_chunks; ++w) v += p[(int64_t)w * stride];
  }
  s_m[jb][c] = v;
  __syncthreads();
  if (j < S.J) {
    double w = 0.;
#pragma unroll
    for (int d = 0; d < LQ_ND; ++d) w = fma(LQ_CT[c * LQ_ND + d], s_m[jb][d], w);
    weights[S.w_first + (int64_t)j * LQ_ND + c] = w;
  }
}

// Rows: the closed form of every row at the nodes of its stream, dotted with the weights.
// One CTA per row; the row's ItemDev carries the stream index in wu_first.
// FAMILY: ITEM_UPS_EXP, ITEM_UPS_GAUSS, ITEM_UPS_EG, ITEM_UPS_GE, ITEM_PHI_EXP, ITEM_PHI_GAUSS
#define LQ_QT 128
template <int FAMILY>
__global__ void __launch_bounds__(LQ_QT)
lag_quad_kernel(const QStreamDev* __restrict__ streams, const ItemDev* __restrict__ rows,
                const double* __restrict__ weights, double* __restrict__ raw) {
  __shared__ double s_red[(LQ_QT / 32) * ASLSD_MAX_COMP];
  const ItemDev it = rows[blockIdx.x];
  const QStreamDev S = streams[it.wu_first];
  const int mask = it.mask;
  UpsGaussConst KG;
  UpsExpConst KE;
  PhiConst KP;
  UpsMixConst KM;
  if (FAMILY == ITEM_UPS_GAUSS) KG.init(it);
  if (FAMILY == ITEM_UPS_EXP) KE.init(it);
  if (FAMILY == ITEM_UPS_EG) KM.init_eg(it);
  if (FAMILY == ITEM_UPS_GE) KM.init_ge(it);
  if (FAMILY == ITEM_PHI_EXP || FAMILY == ITEM_PHI_GAUSS) KP.init(it);
  // bins that meet the row's lag window
  int j_lo = (int)floor(it.s_lo * S.inv_h) - 1, j_hi = (int)floor(it.s_hi * S.inv_h) + 1;
  j_lo = max(j_lo, 0);
  j_hi = min(j_hi, S.J - 1);
  const double* W = weights + S.w_first;
  double acc[ASLSD_MAX_COMP];
#pragma unroll
  for (int c = 0; c < ASLSD_MAX_COMP; ++c) acc[c] = 0.;
  for (int q = j_lo * LQ_ND + threadIdx.x; q < (j_hi + 1) * LQ_ND; q += LQ_QT) {
    const double w = W[q];
    if (w == 0.) continue;
    const int j = q / LQ_ND, i = q - j * LQ_ND;
    const double s = S.h * ((double)j + 0.5 + 0.5 * LQ_X[i]);
    double v[ASLSD_MAX_COMP];
#pragma unroll
    for (int c = 0; c < ASLSD_MAX_COMP; ++c) v[c] = 0.;
    if (FAMILY == ITEM_UPS_GAUSS) eval_ups_gauss<true>(KG, mask, 0., s, v);
    else if (FAMILY == ITEM_UPS_EXP) eval_ups_exp<true>(KE, mask, 0., s, v);
    else if (FAMILY == ITEM_UPS_EG) eval_ups_eg<true>(KM, mask, 0., s, v);
    else if (FAMILY == ITEM_UPS_GE) eval_ups_ge<true>(KM, mask, 0., s, v);
    else if (FAMILY == ITEM_PHI_EXP) eval_phi_exp(KP, mask, s, v);
    else eval_phi_gauss(KP, mask, s, v);
#pragma unroll
    for (int c = 0; c < ASLSD_MAX_COMP; ++c) acc[c] = fma(w, v[c], acc[c]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < ASLSD_MAX_COMP; ++c) {
    double v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp * ASLSD_MAX_COMP + c] = v;
  }
  __syncthreads();
  if (threadIdx.x < ASLSD_MAX_COMP) {
    double v = 0.;
#pragma unroll
    for (int w = 0; w < LQ_QT / 32; ++w) v += s_red[w * ASLSD_MAX_COMP + threadIdx.x];
    // every row owns its destination (the boundary targets of the same row are added by
    // finalize_kernel, which runs before on the same stream)
    raw[it.dst_row * ASLSD_MAX_COMP + threadIdx.x] += v;
  }
}

// Targets near T_f (T_f - t_m below the stream's t_int): the closed forms depend on (T_f - t_m, s),
// so these few targets are evaluated pair by pair - but once per stream, not once per row: one CTA
// = (stream, family) stages the (lag, time-to-end) pairs of the boundary targets in shared memory
// and every warp walks rows over them.  Each row is owned by one warp of one CTA: plain
// read-modify-write of its raw components, fixed order.
#define LB_NT 256
#define LB_PCAP 1536
struct BoundaryGroupDev {
  int32_t stream, row_first, row_count, pad;
  int64_t m_lo, m_hi;      // boundary targets
};

template <int FAMILY>
__global__ void __launch_bounds__(LB_NT)
lag_boundary_kernel(const DimDev* __restrict__ dims, const QStreamDev* __restrict__ streams,
                    const BoundaryGroupDev* __restrict__ groups, const ItemDev* __restrict__ rows, double T_f,
                    double* __restrict__ raw, unsigned long long* __restrict__ counts,
                    const double* __restrict__ weights) {
  __shared__ double s_s[LB_PCAP];
  __shared__ double s_tt[LB_PCAP];
  __shared__ double s_w[LB_PCAP];
  __shared__ int64_t s_warp[LB_NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const BoundaryGroupDev G = groups[blockIdx.x];
  const QStreamDev S = streams[G.stream];
  const DimDev Dt = dims[S.tgt_dim], Ds = dims[S.src_dim];
  if (*Dt.bad) return;   // unsorted upload (aslsd_set_events): the verdict reaches the host later
  const bool same = S.same != 0;
  unsigned long long n_eval = 0;
  for (int64_t m0 = G.m_lo; m0 < G.m_hi; m0 += LB_NT) {
    const int64_t m = m0 + tid;
    double t_m = 0.;
    int64_t lo = 0;
    int32_t cnt = 0;
    if (m < G.m_hi) {
      t_m = Dt.t[m];
      const int64_t hi = same ? m : dim_lower_bound(Ds, t_m);
      lo = min(dim_lower_bound(Ds, t_m - S.s_max), hi);
      cnt = (int32_t)(hi - lo);
    }
    int64_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    __syncthreads();     // previous tile's readers of s_warp / pair buffers are done
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int64_t pre = incl - cnt, total = 0;
#pragma unroll
    for (int w = 0; w < LB_NT / 32; ++w) {
      const int64_t v = s_warp[w];
      if (w < warp) pre += v;
      total += v;
    }
    for (int64_t c0 = 0; c0 < total; c0 += LB_PCAP) {
      const int32_t n_c = (int32_t)min((int64_t)LB_PCAP, total - c0);
      const int32_t ra = (int32_t)min(max(c0 - pre, (int64_t)0), (int64_t)cnt);
      const int32_t rb = (int32_t)min(max(c0 + n_c - pre, (int64_t)0), (int64_t)cnt);
      const bool weighted = S.wa_off >= 0 || S.wb_off >= 0;
      const double w_m = (S.wa_off >= 0 && m < G.m_hi) ? weights[S.wa_off + m] : 1.;
      for (int32_t r = ra; r < rb; ++r) {
        s_s[pre - c0 + r] = t_m - Ds.t[lo + r];
        s_tt[pre - c0 + r] = T_f - t_m;
        if (weighted) s_w[pre - c0 + r] = S.wb_off >= 0 ? w_m * weights[S.wb_off + lo + r] : w_m;
      }
      __syncthreads();
      for (int row = warp; row < G.row_count; row += LB_NT / 32) {
        const ItemDev it = rows[G.row_first + row];
        const int mask = it.mask;
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
        for (int32_t p = lane; p < n_c; p += 32) {
          const double s = s_s[p];
          if (s < it.s_lo || s > it.s_hi) continue;   // the row's own lag window
          const double t = s_tt[p];
          double v[ASLSD_MAX_COMP];
#pragma unroll
          for (int c = 0; c < ASLSD_MAX_COMP; ++c) v[c] = 0.;
          if (FAMILY == ITEM_UPS_GAUSS) eval_ups_gauss<false>(KG, mask, t, s, v);
          else if (FAMILY == ITEM_UPS_EXP) eval_ups_exp<false>(KE, mask, t, s, v);
          else if (FAMILY == ITEM_UPS_EG) eval_ups_eg<false>(KM, mask, t, s, v);
          else if (FAMILY == ITEM_UPS_GE) eval_ups_ge<false>(KM, mask, t, s, v);
          else if (FAMILY == ITEM_PHI_EXP) eval_phi_exp(KP, mask, s, v);
          else eval_phi_gauss(KP, mask, s, v);
          const double w = weighted ? s_w[p] : 1.;
#pragma unroll
          for (int c = 0; c < ASLSD_MAX_COMP; ++c) acc[c] = fma(w, v[c], acc[c]);
          ++n_eval;
        }
#pragma unroll
        for (int c = 0; c < ASLSD_MAX_COMP; ++c) {
          double v = acc[c];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          acc[c] = v;
        }
        if (lane < ASLSD_MAX_COMP) {
          double v = 0.;
#pragma unroll
          for (int c = 0; c < ASLSD_MAX_COMP; ++c) v = (lane == c) ? acc[c] : v;
          raw[it.dst_row * ASLSD_MAX_COMP + lane] += v;
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_eval += __shfl_xor_sync(0xffffffffu, n_eval, o);
  __shared__ unsigned long long s_cnt[LB_NT / 32];
  if (lane == 0) s_cnt[warp] = n_eval;
  __syncthreads();
  if (tid == 0) {
    unsigned long long c = 0;
    for (int w = 0; w < LB_NT / 32; ++w) c += s_cnt[w];
    counts[blockIdx.x] = c;
  }
}
