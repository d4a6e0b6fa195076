// Sampled-lag sums of the stratified (ASLSD) estimator and the fused
// stratum-variance / allocation update.  Replaces the bodies of
// MHPStratEstim.estimate_sum_{adaptive,nonadaptive,exact}dom
// (optimize/estimators/mhp_stratified_estimator.py:534-900) and the index ->
// time transforms of its compute_*_sum closures (:1097-1100, :1148-1151,
// :1196-1203, :1245-1252, :1300-1307).
//
// The (m, n) samples come from the host (the reference's own numpy RNG code,
// utilities/useful_statistics.py:38-127) and are consumed unchanged, so the
// stochastic estimates are those of the reference.  One launch evaluates every
// stratum of every term of one compute_objective call: one CTA per stratum,
// one thread per sample (strided).  Per sample the thread evaluates the term's
// work items (basis / basis-pair closed forms, basis_math.cuh) and applies the
// linear map of plan.h to obtain the function value and every parameter
// derivative; the per-stratum sums are reduced in the CTA.  The allocation
// statistic (f, or ||grad f||_2) of every sample is kept, and two tiny kernels
// turn it into numpy-faithful per-stratum mean / population std (:631-635) and
// into the next relative allocation (:637-645).
#include "basis_math.cuh"
#include "plan.h"

#include <cstring>
#include <limits>

#define SAMP_NT 128
#define SAMP_MAX_OUT 128   // 1 + n_param_1 + n_param_2 of a term

union ItemConst {
  UpsGaussConst g;
  UpsExpConst e;
  PhiConst p;
  UpsMixConst m;
};

struct SampleTermDev {
  int32_t kind, tgt_dim, src_dim, n_out;
  int32_t item_first, item_count;
  double sat_t, v_sat;   // single-index terms: the reference's own value beyond saturation (aslsd_sample_term)
};

struct SampleEntry {
  int32_t comp, c;
  double coef, outer;   // output c += outer * (coef * raw[comp]), each product rounded
};

struct StratumDev {
  int32_t term, alloc_group;   // alloc_group: first stratum of the term's adaptive block (-1: none)
  int64_t begin, end;          // sample range in ix_m / ix_n
  int64_t out_off;             // offset in sum_out
};

__global__ void item_const_kernel(const ItemDev* __restrict__ items, int n_items, ItemConst* __restrict__ consts) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n_items) return;
  const ItemDev it = items[x];
  ItemConst c;
  switch (it.family) {
    case ITEM_UPS_GAUSS: case ITEM_UPS0_GAUSS: c.g.init(it); break;
    case ITEM_UPS_EXP: case ITEM_UPS0_EXP: c.e.init(it); break;
    case ITEM_UPS_EG: case ITEM_UPS0_EG: c.m.init_eg(it); break;
    case ITEM_UPS_GE: case ITEM_UPS0_GE: c.m.init_ge(it); break;
    default: c.p.init(it); break;
  }
  consts[x] = c;
}

// numpy's float64 add.reduce (pairwise summation with 8 partial sums per block of
// at most 128 elements; numpy/core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum)
// restated operation for operation, so that np.mean / np.std of the reference
// (mhp_stratified_estimator.py:631-635) are reproduced bit for bit on
// bit-identical inputs.  The integer sample counts of the next iteration are a
// truncation of the resulting allocations (:587), and in zero-variance strata
// the reference's "standard deviation" is this summation's rounding residue.
// DEV: sum of (a[i] - shift)^2, each step rounded as numpy rounds it (no FMA).
template <bool DEV>
__device__ double np_block_sum(const double* __restrict__ a, int64_t n, double shift) {
  auto f = [&](int64_t i) -> double {
    if (!DEV) return a[i];
    const double dlt = __dsub_rn(a[i], shift);
    return __dmul_rn(dlt, dlt);
  };
  if (n < 8) {
    double res = 0.;
    for (int64_t i = 0; i < n; ++i) res = __dadd_rn(res, f(i));
    return res;
  }
  double r[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) r[q] = f(q);
  int64_t i;
  for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = __dadd_rn(r[q], f(i + q));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, f(i));
  return res;
}

// the recursion sum(a, n) = sum(a, n2) + sum(a + n2, n - n2), n2 = n/2 rounded down
// to a multiple of 8, for n > 128, unrolled on an explicit stack
template <bool DEV>
__device__ double np_pairwise_sum(const double* __restrict__ a, int64_t n, double shift) {
  struct Frame { int64_t off, n; int state; double left; };
  Frame st[64];
  int sp = 0;
  double ret = 0.;
  st[sp++] = Frame{0, n, 0, 0.};
  while (sp > 0) {
    Frame& fr = st[sp - 1];
    if (fr.n <= 128) {
      ret = np_block_sum<DEV>(a + fr.off, fr.n, shift);
      --sp;
      continue;
    }
    int64_t n2 = fr.n / 2;
    n2 -= n2 % 8;
    if (fr.state == 0) {
      fr.state = 1;
      st[sp++] = Frame{fr.off, n2, 0, 0.};
    } else if (fr.state == 1) {
      fr.left = ret;
      fr.state = 2;
      st[sp++] = Frame{fr.off + n2, fr.n - n2, 0, 0.};
    } else {
      ret = __dadd_rn(fr.left, ret);
      --sp;
    }
  }
  return ret;
}

__global__ void __launch_bounds__(SAMP_NT)
sampled_kernel(const DimDev* __restrict__ dims, const ItemDev* __restrict__ items,
               const ItemConst* __restrict__ consts, const int32_t* __restrict__ ent_off,
               const SampleEntry* __restrict__ entries, const SampleTermDev* __restrict__ terms,
               const StratumDev* __restrict__ strata, const int64_t* __restrict__ ix_m,
               const int64_t* __restrict__ ix_n, double T_f, int grad_alloc, double* __restrict__ scratch,
               double* __restrict__ sum_out, const double* __restrict__ weights) {
  __shared__ double s_acc[(SAMP_NT / 32) * SAMP_MAX_OUT];
  const StratumDev S = strata[blockIdx.x];
  const SampleTermDev T = terms[S.term];
  const DimDev Dt = dims[T.tgt_dim], Ds = dims[T.src_dim];
  const int tid = threadIdx.x;
  const int n_out = T.n_out;
  // per-event weights of a marked model (all items of a term share them)
  const int64_t wa_off = T.item_count > 0 ? items[T.item_first].wa_off : -1;
  const int64_t wb_off = T.item_count > 0 ? items[T.item_first].wb_off : -1;
  double acc[SAMP_MAX_OUT], val[SAMP_MAX_OUT];
  for (int c = 0; c < n_out; ++c) acc[c] = 0.;
  for (int64_t x = S.begin + tid; x < S.end; x += SAMP_NT) {
    const int64_t m = ix_m[x];
    const double tm = Dt.t[m];
    double t = T_f - tm, s = 0.;
    if (T.kind >= 2) s = tm - Ds.t[ix_n[x]];
    for (int c = 0; c < n_out; ++c) val[c] = 0.;
    for (int q = 0; q < T.item_count; ++q) {
      const int ix = T.item_first + q;
      const ItemDev& it = items[ix];
      const int mask = it.mask;
      double raw[ASLSD_MAX_COMP];
#pragma unroll
      for (int c = 0; c < ASLSD_MAX_COMP; ++c) raw[c] = 0.;
      const bool interior = t >= it.t_int;
      switch (it.family) {
        case ITEM_UPS_GAUSS: case ITEM_UPS0_GAUSS:
          if (interior) eval_ups_gauss<true>(consts[ix].g, mask, 0., s, raw);
          else eval_ups_gauss<false>(consts[ix].g, mask, t, s, raw);
          break;
        case ITEM_UPS_EXP: case ITEM_UPS0_EXP:
          if (interior) eval_ups_exp<true>(consts[ix].e, mask, 0., s, raw);
          else eval_ups_exp<false>(consts[ix].e, mask, t, s, raw);
          break;
        case ITEM_UPS_EG: case ITEM_UPS0_EG:
          if (interior) eval_ups_eg<true>(consts[ix].m, mask, 0., s, raw);
          else eval_ups_eg<false>(consts[ix].m, mask, t, s, raw);
          break;
        case ITEM_UPS_GE: case ITEM_UPS0_GE:
          if (interior) eval_ups_ge<true>(consts[ix].m, mask, 0., s, raw);
          else eval_ups_ge<false>(consts[ix].m, mask, t, s, raw);
          break;
        case ITEM_PHI_EXP: eval_phi_exp(consts[ix].p, mask, s, raw); break;
        case ITEM_PHI_GAUSS: eval_phi_gauss(consts[ix].p, mask, s, raw); break;
        case ITEM_PSI_EXP: eval_psi_exp(consts[ix].p, mask, t, raw); break;
        default: eval_psi_gauss(consts[ix].p, mask, t, raw); break;
      }
      for (int e = ent_off[ix]; e < ent_off[ix + 1]; ++e) {
        const SampleEntry E = entries[e];
        val[E.c] = __dadd_rn(val[E.c], __dmul_rn(E.outer, __dmul_rn(E.coef, raw[E.comp])));
      }
    }
    // beyond saturation the value is the caller's constant: the reference's own arithmetic
    // (aslsd_sample_term; a NaN threshold compares false)
    if (T.kind < 2 && t >= T.sat_t) val[0] = T.v_sat;
    if (wa_off >= 0 || wb_off >= 0) {
      // f = kernel functional x impact functional, impact product first
      // (mtlh_stratified_estimator.py:1530-1533, :686 np.prod over the two functions)
      double w = wa_off >= 0 ? weights[wa_off + m] : 1.;
      if (wb_off >= 0) w = __dmul_rn(w, weights[wb_off + ix_n[x]]);
      for (int c = 0; c < n_out; ++c) val[c] = __dmul_rn(val[c], w);
    }
    double stat;
    if (grad_alloc) {
      double g2 = 0.;   // grad_norm_vals += diff_f_vals**2, one parameter after the other (:615)
      for (int c = 1; c < n_out; ++c) g2 = __dadd_rn(g2, __dmul_rn(val[c], val[c]));
      stat = sqrt(g2);
    } else {
      stat = val[0];
    }
    scratch[x] = stat;
    for (int c = 0; c < n_out; ++c) acc[c] += val[c];
  }
  // per-stratum sums of the function value and of every derivative
  const int lane = tid & 31, warp = tid >> 5;
  for (int c = 0; c < n_out; ++c) {
    double v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_acc[warp * SAMP_MAX_OUT + c] = v;
  }
  __syncthreads();
  for (int c = tid; c < n_out; c += SAMP_NT) {
    double v = 0.;
#pragma unroll
    for (int w = 0; w < SAMP_NT / 32; ++w) v += s_acc[w * SAMP_MAX_OUT + c];
    sum_out[S.out_off + c] = v;
  }
}

// Allocation statistic of every stratum: numpy.mean and numpy.std (ddof = 0) of
// the per-sample values left in `scratch`, in numpy's own summation order.  One
// thread per stratum (a stratum holds 10^2 - 10^4 samples; strata run in parallel).
__global__ void stat_kernel(const StratumDev* __restrict__ strata, int n_strata,
                            const double* __restrict__ scratch, double* __restrict__ stat_mean,
                            double* __restrict__ stat_std) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_strata) return;
  const StratumDev S = strata[s];
  const int64_t n = S.end - S.begin;
  if (n <= 0) { stat_mean[s] = 0.; stat_std[s] = 0.; return; }
  const double* a = scratch + S.begin;
  const double mean = __ddiv_rn(np_pairwise_sum<false>(a, n, 0.), (double)n);
  const double var = __ddiv_rn(np_pairwise_sum<true>(a, n, mean), (double)n);
  stat_mean[s] = mean;
  stat_std[s] = sqrt(var);
}

// rel_alloc[s] = size_s * sqrt(size_s / (size_s - 1)) * std_s, normalised over the
// group; default allocation when every std is zero (:637-645).  One thread per
// group (groups hold a few tens of strata).
__global__ void allocation_kernel(const int32_t* __restrict__ group_first, const int32_t* __restrict__ group_count,
                                  int n_groups, const double* __restrict__ sizes,
                                  const double* __restrict__ stat_std, const double* __restrict__ default_alloc,
                                  double* __restrict__ rel_alloc) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const int a = group_first[g], n = group_count[g];
  for (int s = a; s < a + n; ++s) {
    const double size = sizes[s];
    // strata_size * np.sqrt(strata_size / (strata_size - 1.)) * std, left to right
    const double fpop = sqrt(__ddiv_rn(size, __dsub_rn(size, 1.)));
    rel_alloc[s] = __dmul_rn(__dmul_rn(size, fpop), stat_std[s]);
  }
  const double tot = np_pairwise_sum<false>(rel_alloc + a, n, 0.);   // np.sum(rel_alloc)
  for (int s = a; s < a + n; ++s) rel_alloc[s] = tot > 0. ? __ddiv_rn(rel_alloc[s], tot) : default_alloc[s];
}

struct Upload {
  // one pinned staging block + one device block per call, carved into arrays
  size_t bytes = 0;
  std::vector<std::pair<const void*, size_t>> parts;
  std::vector<size_t> offs;
  size_t add(const void* p, size_t n) {
    size_t off = (bytes + 15) & ~(size_t)15;
    parts.push_back({p, n});
    offs.push_back(off);
    bytes = off + n;
    return off;
  }
};

extern "C" int aslsd_sampled_n_out(const aslsd_ctx* ctx, int k, int term_kind, int i, int j, int32_t* n_out) {
  if (!ctx || !ctx->has_model || k < 0 || k >= ctx->d || i < 0 || i >= ctx->d || !n_out) return ASLSD_ERR_ARG;
  const int d = ctx->d;
  const int ni = ctx->kernels[k * d + i].n_param();
  switch (term_kind) {
    case 0: *n_out = 1 + ni + 1; return ASLSD_OK;
    case 1: case 2: *n_out = 1 + ni; return ASLSD_OK;
    case 3:
      if (j < 0 || j >= d) return ASLSD_ERR_ARG;
      *n_out = (i == j) ? 1 + ni : 1 + ni + ctx->kernels[k * d + j].n_param();
      return ASLSD_OK;
  }
  return ASLSD_ERR_ARG;
}

extern "C" int aslsd_sampled_sums(aslsd_ctx* ctx, int k, const double* x_k, int flags, int n_terms,
                                  const aslsd_sample_term* terms, const int64_t* stratum_offsets,
                                  const int64_t* ix_m, const int64_t* ix_n, const double* strata_sizes,
                                  const double* default_alloc, double* sum_out, double* stat_mean,
                                  double* stat_std, double* rel_alloc) {
  if (!ctx || !x_k || n_terms < 0 || (n_terms && (!terms || !stratum_offsets || !sum_out)))
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "sampled sums: bad arguments");
  if (n_terms == 0) return ASLSD_OK;
  if (int vrc = events_verdict(ctx, true)) return vrc;
  SumPlan P;
  int rc = build_plan(ctx, k, x_k, flags | ASLSD_F, nullptr, nullptr, true, P);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // ---- host-side tables ----------------------------------------------------------------
  std::vector<int32_t> ent_off(P.items.size() + 1, 0);
  std::vector<std::vector<SampleEntry>> per_item(P.items.size());
  std::vector<SampleTermDev> tdev(n_terms);
  std::vector<StratumDev> sdev;
  std::vector<int32_t> group_first, group_count;
  std::vector<bool> emitted(P.terms.size(), false);
  int64_t out_pos = 0, n_strata = 0;
  for (int q = 0; q < n_terms; ++q) n_strata += terms[q].n_strata;
  const int64_t n_samples = stratum_offsets[n_strata];
  bool need_n = false;
  int64_t s_ix = 0;
  for (int q = 0; q < n_terms; ++q) {
    const aslsd_sample_term& A = terms[q];
    const int tix = find_term(ctx, A.kind, A.i, A.kind == 3 ? A.j : A.i);
    if (tix < 0 || A.n_strata < 0) return aslsd_fail(ctx, ASLSD_ERR_ARG, "sampled sums: unknown term");
    const TermPlan& T = P.terms[tix];
    SampleTermDev& D = tdev[q];
    D.kind = T.kind;
    D.tgt_dim = (T.kind == 2) ? k : T.i;
    D.src_dim = (T.kind == 2) ? T.i : T.j;
    D.n_out = 1 + T.n1 + T.n2;
    D.item_first = T.item_first;
    D.item_count = T.item_count;
    D.sat_t = (T.kind < 2 && A.sat_t >= 0.) ? A.sat_t : std::numeric_limits<double>::infinity();
    D.v_sat = A.v_sat;
    if (D.n_out > SAMP_MAX_OUT)
      return aslsd_fail(ctx, ASLSD_ERR_UNSUPPORTED, "sampled sums: more than 127 parameters in one term");
    if (T.kind >= 2) need_n = true;
    if (!emitted[tix]) {   // a term may appear more than once in a batch
      emitted[tix] = true;
      emit_term(ctx, P, (size_t)tix, x_k, [&](int c, int item, int comp, double coef, double outer) {
        per_item[item].push_back(SampleEntry{comp, c, coef, outer});
      });
    }
    // index validation: an out-of-range sample would fault the device
    const int64_t n_tgt = ctx->n_events[D.tgt_dim], n_src = ctx->n_events[D.src_dim];
    int32_t group = -1;
    for (int s = 0; s < A.n_strata; ++s, ++s_ix) {
      const int64_t a = stratum_offsets[s_ix], b = stratum_offsets[s_ix + 1];
      if (a > b || a < 0) return aslsd_fail(ctx, ASLSD_ERR_ARG, "sampled sums: stratum offsets must be non-decreasing");
      for (int64_t x = a; x < b; ++x) {
        if (ix_m[x] < 0 || ix_m[x] >= n_tgt)
          return aslsd_fail(ctx, ASLSD_ERR_ARG, "sampled sums: target index out of range");
        if (T.kind >= 2 && (!ix_n || ix_n[x] < 0 || ix_n[x] >= n_src))
          return aslsd_fail(ctx, ASLSD_ERR_ARG, "sampled sums: source index out of range");
      }
      StratumDev S;
      S.term = q;
      S.begin = a;
      S.end = b;
      S.out_off = out_pos;
      out_pos += D.n_out;
      // consecutive strata of a term with a positive size form one allocation group
      const bool adaptive = rel_alloc && strata_sizes && strata_sizes[s_ix] > 0.;
      if (adaptive) {
        if (group < 0) {
          group = (int32_t)s_ix;
          group_first.push_back(group);
          group_count.push_back(0);
        }
        group_count.back()++;
      } else {
        group = -1;
      }
      S.alloc_group = adaptive ? group : -1;
      sdev.push_back(S);
    }
  }
  std::vector<SampleEntry> entries;
  for (size_t it = 0; it < per_item.size(); ++it) {
    ent_off[it] = (int32_t)entries.size();
    entries.insert(entries.end(), per_item[it].begin(), per_item[it].end());
  }
  ent_off[per_item.size()] = (int32_t)entries.size();
  if (entries.empty()) entries.push_back(SampleEntry{0, 0, 0., 0.});
  const size_t n_items = P.items.size();
  const size_t n_groups = group_first.size();

  // ---- one staging block up, one result block down -------------------------------------------
  Upload up;
  const size_t o_items = up.add(P.items.data(), n_items * sizeof(ItemDev));
  const size_t o_entoff = up.add(ent_off.data(), ent_off.size() * sizeof(int32_t));
  const size_t o_ent = up.add(entries.data(), entries.size() * sizeof(SampleEntry));
  const size_t o_terms = up.add(tdev.data(), tdev.size() * sizeof(SampleTermDev));
  const size_t o_strata = up.add(sdev.data(), sdev.size() * sizeof(StratumDev));
  const size_t o_m = up.add(ix_m, (size_t)n_samples * sizeof(int64_t));
  const size_t o_n = need_n ? up.add(ix_n, (size_t)n_samples * sizeof(int64_t)) : 0;
  const size_t o_sizes = n_groups ? up.add(strata_sizes, (size_t)n_strata * sizeof(double)) : 0;
  const size_t o_def = n_groups ? up.add(default_alloc, (size_t)n_strata * sizeof(double)) : 0;
  const size_t o_gf = n_groups ? up.add(group_first.data(), n_groups * sizeof(int32_t)) : 0;
  const size_t o_gc = n_groups ? up.add(group_count.data(), n_groups * sizeof(int32_t)) : 0;
  // device-only regions
  auto reserve = [&](size_t n) { size_t off = (up.bytes + 15) & ~(size_t)15; up.bytes = off + n; return off; };
  const size_t o_consts = reserve(n_items * sizeof(ItemConst));
  const size_t o_scratch = reserve((size_t)std::max<int64_t>(n_samples, 1) * sizeof(double));
  const size_t o_res = reserve(0);
  const size_t r_sum = 0;
  const size_t r_mean = r_sum + (size_t)out_pos * sizeof(double);
  const size_t r_std = r_mean + (size_t)n_strata * sizeof(double);
  const size_t r_alloc = r_std + (size_t)n_strata * sizeof(double);
  const size_t res_bytes = r_alloc + (size_t)n_strata * sizeof(double);
  up.bytes = o_res + res_bytes;
  if ((rc = ensure_device_buffer(ctx, &ctx->d_sample, &ctx->sample_cap, up.bytes))) return rc;
  if ((rc = ensure_pinned_buffer(ctx, std::max(o_consts, res_bytes)))) return rc;
  char* hp = (char*)ctx->h_pinned;
  for (size_t a = 0; a < up.parts.size(); ++a)
    if (up.parts[a].second) std::memcpy(hp + up.offs[a], up.parts[a].first, up.parts[a].second);
  char* dp = (char*)ctx->d_sample;
  CUDA_TRY(ctx, cudaMemcpyAsync(dp, hp, o_consts, cudaMemcpyHostToDevice, ctx->stream));
  ctx->last_launches = 0;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  if (n_items) {
    item_const_kernel<<<(unsigned)((n_items + 127) / 128), 128, 0, ctx->stream>>>(
        (const ItemDev*)(dp + o_items), (int)n_items, (ItemConst*)(dp + o_consts));
    ctx->last_launches++;
  }
  double* d_res = (double*)(dp + o_res);
  if (n_strata) {
    sampled_kernel<<<(unsigned)n_strata, SAMP_NT, 0, ctx->stream>>>(
        (const DimDev*)ctx->d_dims, (const ItemDev*)(dp + o_items), (const ItemConst*)(dp + o_consts),
        (const int32_t*)(dp + o_entoff), (const SampleEntry*)(dp + o_ent), (const SampleTermDev*)(dp + o_terms),
        (const StratumDev*)(dp + o_strata), (const int64_t*)(dp + o_m),
        need_n ? (const int64_t*)(dp + o_n) : nullptr, ctx->T_f, (flags & ASLSD_GRAD_ALLOC) ? 1 : 0,
        (double*)(dp + o_scratch), d_res + r_sum / sizeof(double), ctx->d_weights);
    ctx->last_launches++;
    stat_kernel<<<(unsigned)((n_strata + 63) / 64), 64, 0, ctx->stream>>>(
        (const StratumDev*)(dp + o_strata), (int)n_strata, (const double*)(dp + o_scratch),
        d_res + r_mean / sizeof(double), d_res + r_std / sizeof(double));
    ctx->last_launches++;
  }
  if (n_groups) {
    allocation_kernel<<<(unsigned)((n_groups + 63) / 64), 64, 0, ctx->stream>>>(
        (const int32_t*)(dp + o_gf), (const int32_t*)(dp + o_gc), (int)n_groups, (const double*)(dp + o_sizes),
        d_res + r_std / sizeof(double), (const double*)(dp + o_def), d_res + r_alloc / sizeof(double));
    ctx->last_launches++;
  }
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timing_pending = true;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(hp, d_res, res_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  std::memcpy(sum_out, hp + r_sum, (size_t)out_pos * sizeof(double));
  if (stat_mean) std::memcpy(stat_mean, hp + r_mean, (size_t)n_strata * sizeof(double));
  if (stat_std) std::memcpy(stat_std, hp + r_std, (size_t)n_strata * sizeof(double));
  if (rel_alloc) {
    const double* ra = (const double*)(hp + r_alloc);
    for (int64_t s = 0; s < n_strata; ++s)
      rel_alloc[s] = (sdev[s].alloc_group >= 0) ? ra[s] : (default_alloc ? default_alloc[s] : 0.);
  }
  return ASLSD_OK;
}

// Marked models: the sampled sums with per-event weights (roles as in
// aslsd_exact_sums_weighted); replaces the products with the impact functional in
// MTLHStratEstim.estimate_sum_*dom (optimize/estimators/mtlh_stratified_estimator.py:615-1025).
extern "C" int aslsd_sampled_sums_weighted(aslsd_ctx* ctx, int k, const double* x_k, int flags, int n_terms,
                                           const aslsd_sample_term* terms, const int64_t* stratum_offsets,
                                           const int64_t* ix_m, const int64_t* ix_n, const double* strata_sizes,
                                           const double* default_alloc, const int32_t* roles, double* sum_out,
                                           double* stat_mean, double* stat_std, double* rel_alloc) {
  if (!ctx || !roles) return aslsd_fail(ctx, ASLSD_ERR_ARG, "sampled_sums_weighted: bad arguments");
  const int d = ctx->d;
  for (int r = 0; r < 5; ++r)
    for (int i = 0; i < d; ++i) {
      const int v = roles[r * d + i];
      if (v >= 0 && (v >= (int)ctx->weight_off.size() || ctx->weight_dim[v] != i))
        return aslsd_fail(ctx, ASLSD_ERR_ARG, "sampled_sums_weighted: weight vector of another event type");
    }
  ctx->roles = roles;
  const int rc = aslsd_sampled_sums(ctx, k, x_k, flags, n_terms, terms, stratum_offsets, ix_m, ix_n, strata_sizes,
                                    default_alloc, sum_out, stat_mean, stat_std, rel_alloc);
  ctx->roles = nullptr;
  return rc;
}

// Stand-alone form of the allocation update for callers that keep their own
// statistics (same kernel, one group).
extern "C" int aslsd_update_allocation(aslsd_ctx* ctx, int n_strata, const double* strata_sizes,
                                       const double* stat_std, const double* default_alloc,
                                       double* rel_alloc_out) {
  if (!ctx || n_strata <= 0 || !strata_sizes || !stat_std || !default_alloc || !rel_alloc_out)
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "allocation update: bad arguments");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)n_strata;
  const size_t bytes = 4 * n * sizeof(double) + 2 * sizeof(int32_t) + 16;
  int rc;
  if ((rc = ensure_device_buffer(ctx, &ctx->d_sample, &ctx->sample_cap, bytes))) return rc;
  if ((rc = ensure_pinned_buffer(ctx, bytes))) return rc;
  double* hp = (double*)ctx->h_pinned;
  std::memcpy(hp, strata_sizes, n * sizeof(double));
  std::memcpy(hp + n, stat_std, n * sizeof(double));
  std::memcpy(hp + 2 * n, default_alloc, n * sizeof(double));
  int32_t* hg = (int32_t*)(hp + 4 * n);
  hg[0] = 0;
  hg[1] = n_strata;
  double* dp = (double*)ctx->d_sample;
  CUDA_TRY(ctx, cudaMemcpyAsync(dp, hp, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const int32_t* dg = (const int32_t*)(dp + 4 * n);
  allocation_kernel<<<1, 64, 0, ctx->stream>>>(dg, dg + 1, 1, dp, dp + n, dp + 2 * n, dp + 3 * n);
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(hp + 3 * n, dp + 3 * n, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  std::memcpy(rel_alloc_out, hp + 3 * n, n * sizeof(double));
  return ASLSD_OK;
}
