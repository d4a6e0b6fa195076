// Exact LSE sums: batching over sub-problems, de-duplication of identical work
// items, lag-quadrature streams, launches and assembly.
// Replaces MHPExactEstim.compute_objective and its five compute_*_sum methods
// (optimize/estimators/mhp_exact_estimator.py:543-761) for one or several k.
//
// A batch is a list of sub-problems (k, x_k).  Every sub-problem is planned into
// work items (plan.cu).  The raw components of an item depend only on (family,
// event streams, target block, beta/delta of its bases, component mask) - not on
// omega, not on k - so identical items of different terms / sub-problems are
// evaluated once ("rows").  For the SBF models of the reference (beta, delta
// fixed) every upsilon_i_j row is shared by all d sub-problems.
//
// For the interior targets (T_f - t_m beyond the kernels' reach) every double-sum row
// is a function of the lag alone: all rows of one ordered pair of event streams
// share one pass over the event pairs (lag_quadrature.cuh).  The last few targets
// of every row, and streams that do not qualify, go through the per-item pair
// kernels (pair_kernels.cuh).  Single-index rows (psi, upsilon at lag zero) are
// constant in the interior: count x closed form, per-target evaluation only near T_f.
#include "lag_quadrature.cuh"
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <thread>

struct RowKey {
  ItemDev it;
  bool operator<(const RowKey& o) const { return std::memcmp(&it, &o.it, sizeof(ItemDev)) < 0; }
};

struct BatchPlan {
  std::vector<double> key_rows, key_work;   // cache keys (see get_batch)
  bool has_work = false;
  bool device_resident = false;   // item / stream tables of this plan are on the device
  bool weights_valid = false;     // the context's node weights were computed for this work part
  int flags = 0;
  std::vector<int> ks;
  std::vector<SumPlan> plans;
  std::vector<std::vector<int64_t>> item_row;   // per plan: item -> row
  std::vector<ItemDev> rows;                    // canonical item of every row
  int64_t n_items_total = 0;                    // work items before de-duplication
  std::vector<int64_t> lo, hi;                  // target block per dimension (build_work)
  // per-item kernels
  std::vector<ItemDev> items;                   // device items, grouped by family
  int fam_first[ITEM_N_FAMILIES], fam_count[ITEM_N_FAMILIES];
  int64_t fam_wu_first[ITEM_N_FAMILIES], fam_wu_count[ITEM_N_FAMILIES];
  int64_t n_wu = 0;
  // lag quadrature
  std::vector<QStreamDev> qstreams;
  std::vector<ItemDev> qrows;                   // grouped by family; wu_first = stream index
  int qfam_first[ITEM_N_FAMILIES], qfam_count[ITEM_N_FAMILIES];
  std::vector<BoundaryGroupDev> bgroups;        // targets near T_f of the quadrature rows, by family
  int bfam_first[ITEM_N_FAMILIES], bfam_count[ITEM_N_FAMILIES];
  int64_t n_wu_q = 0, n_q_partials = 0, n_q_weights = 0, n_q_evals = 0;
  // streams are ordered by category - unweighted streams of the small CTA shape, unweighted streams
  // of the big one, weighted streams - and each category is one launch over its work units
  int64_t q_cat_wu[4] = {0, 0, 0, 0};           // first work unit of every category (prefix)
  size_t q_cat_smem[3] = {0, 0, 0};
  // work part in two phases: the streams first (all the moment kernels need), the rows while those run
  std::vector<std::vector<int64_t>> q_groups;   // rows of every (streams, target block, weights) group, launch order
  std::vector<int> q_group_stream;              // its stream (index in qstreams), -1: per-item kernels only
  bool rows_pending = false;                    // build_work_rows has not run for the current streams
  // streams of the small CTA shape are ordered by the later of their two event types: the work units of
  // [q_arrival_wu[a], q_arrival_wu[a + 1]) need the types 0..a only, and can start while the other
  // types are still being uploaded
  std::vector<int64_t> q_arrival_wu;
  // scratch of build_work, kept between calls (no reallocation when only the events change)
  std::vector<ItemDev> scratch_q[ITEM_N_FAMILIES], scratch_f[ITEM_N_FAMILIES];
  int q_jmax = 0;
  // device assembly (raw rows -> packed sums): the linear map in CSR form, built once per row plan
  std::vector<int64_t> out_off;                 // per sub-problem: offset in the concatenated packed sums
  int64_t out_total = 0;
  int w_stride = 0;                             // omega table entries per sub-problem
  bool asm_built = false;
  int64_t asm_entries = 0;
  void* d_asm = nullptr;                        // AsmEntry[asm_entries], owned by the plan
  int64_t* d_asm_ptr = nullptr;                 // out_total + 1 row starts
  int uses = 0;                                 // calls served by this row plan
  int64_t uid = 0;                              // distinguishes a plan from an earlier one at the same address
  ~BatchPlan() {
    if (d_asm) cudaFree(d_asm);
    if (d_asm_ptr) cudaFree(d_asm_ptr);
  }
};

// The last few batch plans, most recent first (one per sub-problem when a fit loop takes the
// sub-problems in turn, plus the batched one).
struct PlanCache {
  std::vector<BatchPlan*> plans;
  ~PlanCache() { for (auto* p : plans) delete p; }
};
#define PLAN_CACHE_MAX 24

// One term of the assembly map: out += w[o] * ((((w[a] * w[b]) * g) / gd) * raw[raw_ix])
struct AsmEntry {
  int32_t raw_ix, a, b, o;
  double g, gd;
};

static ItemDev canonical(const ItemDev& it) {
  ItemDev c = it;
  c.n_chunks = 0;
  c.tpw = 0;
  c.wu_first = 0;
  c.dst_row = 0;
  return c;
}

// first target index in [m_lo, m_hi) whose time to the end is below t_int
static int64_t interior_end(const aslsd_ctx* ctx, int dim, int64_t m_lo, int64_t m_hi, double t_int) {
  const double* t = ctx->h_times[dim];
  int64_t lo = m_lo, hi = m_hi;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ctx->T_f - t[mid] >= t_int) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static bool is_pair_family(int f) { return item_pair_family(f); }

// Smallest length scale (in the lag) of the functions a row sums: the bin width of
// its stream is a fixed fraction of it (lag_quadrature.cuh).
static double row_scale(const ItemDev& it) {
  const bool own = (it.mask & (MASK(C_SW) | MASK(C_SB) | MASK(C_SD))) != 0;
  switch (it.family) {
    case ITEM_PHI_EXP: return 1. / it.b1;
    case ITEM_PHI_GAUSS: return it.b1;
    case ITEM_UPS_EXP: return 1. / (own ? std::max(it.b2, it.bo) : it.b2);
    // e^{beta1 s} against a Gaussian tail of std beta2 / e^{-beta2 s} against limits of slope 1 / beta1
    // (0.7: the product of the two factors varies faster than either one)
    case ITEM_UPS_EG: return 0.7 * std::min(1. / it.b1, it.b2);
    case ITEM_UPS_GE: return 0.7 * std::min(1. / it.b2, it.b1);
    default: {
      // P(s): std sigma; lower limit of the integral: slope beta1 / (beta2 sigma) in s;
      // own-basis forms: std sqrt(2) beta
      const double sig = std::sqrt(it.b1 * it.b1 + it.b2 * it.b2);
      double sc = std::min(sig, it.b2 * sig / it.b1);
      if (own) sc = std::min(sc, std::sqrt(2.) * it.bo);
      return sc;
    }
  }
}

// Unweighted streams of at most 64 bins: one warp per work unit (default) or the 128-thread CTAs
static bool lq_warp_shape_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("ASLSD_LQ_WARP");
    return e == nullptr ? (LQW_DEFAULT != 0) : std::atoi(e) != 0;
  }();
  return on;
}

// Turn the pair rows of one (target stream, source stream, target block) into a
// quadrature stream, if the geometry allows it.
// Returns its index in B.qstreams, or -1.
static int build_stream(const aslsd_ctx* ctx, BatchPlan& B, const std::vector<int64_t>& group) {
  const ItemDev first = B.rows[group[0]];
  double t_int = 0., s_max = 0., scale = 1e300;
  for (int64_t r : group) {
    const ItemDev& it = B.rows[r];
    const bool is_phi = it.family == ITEM_PHI_EXP || it.family == ITEM_PHI_GAUSS;
    if (!is_phi) t_int = std::max(t_int, it.t_int);   // phi does not depend on T_f - t_m
    s_max = std::max(s_max, it.s_hi);
    scale = std::min(scale, row_scale(it));
  }
  if (!(scale > 0.) || !(s_max > 0.) || !std::isfinite(s_max) || !std::isfinite(t_int)) return -1;
  // the moment kernel counts the pairs of a tile (at most 512 targets x their windows) in 32 bits: a
  // reach that could hold ~2^31 / 512 sources per target (16 x the mean density as margin for
  // clustering) stays with the per-item kernels, which count in 64 bits
  {
    const int64_t n_src = ctx->n_events[first.src_dim];
    const double* ts = ctx->h_times[first.src_dim];
    const double span = n_src > 1 ? ts[n_src - 1] - ts[0] : 0.;
    const double per_target = span > 0. ? std::min<double>((double)n_src, 16. * (double)n_src * s_max / span)
                                        : (double)n_src;
    if (per_target * (double)LqBig::TM >= 2.0e9) return -1;
  }
  const int64_t m_b = interior_end(ctx, first.tgt_dim, first.m_lo, first.m_hi, t_int);
  if (m_b - first.m_lo < LqBig::TM / 2) return -1;   // the per-item kernels take it all
  // bins [j h, (j + 1) h) up to s_max plus a spare one (a lag can exceed s_max by a rounding
  // error); the count is rounded up to a multiple of 32 - the moment kernel hands bins to
  // groups of lanes in rounds of 64 - and the bins shrink accordingly
  const double h_max = LQ_SCALE_FRACTION * scale;
  double nb = std::floor(s_max / h_max) + 2.;
  if (!(nb <= (double)LQ_JMAX)) return -1;
  nb = std::min<double>(LQ_JMAX, 32. * std::ceil(nb / 32.));
  const double h = s_max / (nb - 1.5);
  const bool weighted = first.wa_off >= 0 || first.wb_off >= 0;
  // CTA shape: the small one where every pair of its lanes can own a bin for a whole work unit
  const bool small = !weighted && nb <= (double)LqSmall::ONE_BIN_J;
  // ... or, by default, one warp per work unit (lag_moment_warp_kernel; ASLSD_LQ_WARP=0: the small CTA shape)
  const bool warp_shape = small && lq_warp_shape_enabled();
  const int ucap = warp_shape ? LQW_PC
                              : (small ? lq_pick_ucap<LqSmall>((int)nb, false) : lq_pick_ucap<LqBig>((int)nb, weighted));
  if (ucap == 0) return -1;
  QStreamDev S;
  std::memset(&S, 0, sizeof(S));
  S.tgt_dim = first.tgt_dim;
  S.src_dim = first.src_dim;
  S.same = first.same;
  S.J = (int32_t)nb;
  S.ucap = ucap;
  S.shape = warp_shape ? LQ_SHAPE_WARP : (small ? LQ_SHAPE_SMALL : LQ_SHAPE_BIG);
  if (warp_shape) S.tm = LQW_TM;
  else {
    // targets per tile: the pairs of a tile should fit one chunk (a second, nearly empty chunk pays the
    // barriers, scans and short loops of a full one) - 90 % of a chunk at the stream's mean density.
    // C5 (14 sources per target, 1 792 pairs per chunk, 128 targets overflowing every second tile):
    // 10.56 ms at 128 targets, 10.45 at 115, 10.74 at 102, 11.97 at 77 - the fixed cost per tile
    // (searches, scans, barriers) weighs more than the second chunk (gpurun_out / profiles r2_o_fill)
    const int nt_shape = small ? LqSmall::NT : LqBig::NT;
    const int64_t n_src = ctx->n_events[first.src_dim];
    const double* ts = ctx->h_times[first.src_dim];
    const double span = n_src > 1 ? ts[n_src - 1] - ts[0] : 0.;
    const double per_target = span > 0. ? std::max(1e-3, (double)n_src * s_max / span) : (double)n_src;
    double tmf = 0.9 * (double)ucap / per_target;
    if (const char* e = std::getenv("ASLSD_LQ_FILL")) tmf = std::atof(e) * (double)ucap / per_target;   // tuning
    S.tm = (int32_t)std::max(8., std::min((double)nt_shape, std::floor(tmf)));
  }
  S.wa_off = first.wa_off;
  S.wb_off = first.wb_off;
  while ((1 << S.nbits) < S.J) ++S.nbits;
  S.m_lo = first.m_lo;
  S.m_hi = m_b;
  S.s_max = s_max;
  S.h = h;
  S.inv_h = 1. / h;
  const int sx = (int)B.qstreams.size();
  B.qstreams.push_back(S);
  B.q_jmax = std::max(B.q_jmax, (int)S.J);
  {
    const int cat = weighted ? 2 : (small ? 0 : 1);
    const size_t bytes = warp_shape ? 0 : (small ? lq_smem_bytes<LqSmall>(S.J, S.ucap, false)
                                                 : lq_smem_bytes<LqBig>(S.J, S.ucap, weighted));
    B.q_cat_smem[cat] = std::max(B.q_cat_smem[cat], bytes);
  }
  return sx;
}

// Rows: the work items of every sub-problem and their de-duplication.  Pure function of
// the model, the flags and the non-omega parameters (target ranges are patched in by
// build_work), so it survives a change of the event streams.
static int build_rows(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs, int flags,
                      BatchPlan& B) {
  B.flags = flags;
  B.ks.assign(ks, ks + n_k);
  B.plans.resize(n_k);
  B.item_row.resize(n_k);
  std::map<RowKey, int64_t> index;
  for (int q = 0; q < n_k; ++q) {
    int rc = build_plan(ctx, ks[q], xs[q], flags, nullptr, nullptr, false, B.plans[q]);
    if (rc) return rc;
    const SumPlan& P = B.plans[q];
    B.item_row[q].resize(P.items.size());
    B.n_items_total += (int64_t)P.items.size();
    for (size_t x = 0; x < P.items.size(); ++x) {
      RowKey key;
      key.it = canonical(P.items[x]);
      auto found = index.find(key);
      if (found == index.end()) {
        found = index.emplace(key, (int64_t)B.rows.size()).first;
        B.rows.push_back(key.it);
      }
      B.item_row[q][x] = found->second;
    }
  }
  return ASLSD_OK;
}

// Work: target ranges, quadrature streams, per-item work units.  Depends on the events.  Two phases:
// build_work_streams leaves the stream records complete (all the moment kernels read), build_work_rows
// the rows, boundary groups and per-item work units - run_batch runs it while the moment kernels are
// on the device (the 0.5 - 0.8 ms it takes at C5 would otherwise sit between the upload and the first launch).
static int build_work_streams(aslsd_ctx* ctx, const int64_t* m_lo, const int64_t* m_hi, BatchPlan& B) {
  if (!ctx->d_times || (int)ctx->n_events.size() != ctx->d)
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: no events set");
  B.items.clear();
  B.qstreams.clear();
  B.qrows.clear();
  B.bgroups.clear();
  B.q_groups.clear();
  B.q_group_stream.clear();
  B.n_wu = B.n_wu_q = B.n_q_partials = B.n_q_weights = B.n_q_evals = 0;
  B.q_jmax = 0;
  for (int c = 0; c < 3; ++c) B.q_cat_smem[c] = 0;
  B.device_resident = false;
  B.weights_valid = false;
  B.rows_pending = true;
  const int d = ctx->d;
  B.lo.resize(d);
  B.hi.resize(d);
  for (int i = 0; i < d; ++i) {
    B.lo[i] = m_lo ? std::max<int64_t>(0, m_lo[i]) : 0;
    B.hi[i] = m_hi ? std::min<int64_t>(ctx->n_events[i], m_hi[i]) : ctx->n_events[i];
  }
  for (auto& it : B.rows) {
    it.m_lo = B.lo[it.tgt_dim];
    it.m_hi = B.hi[it.tgt_dim];
  }
  const size_t n_rows = B.rows.size();
  if (ctx->use_quadrature) {
    // group the pair rows by (streams, target block), in row order
    // (marked models: also by the pair of weight vectors - weighted moments)
    std::map<std::array<int64_t, 6>, size_t> groups;
    std::vector<std::vector<int64_t>> found;
    std::vector<bool> weighted;
    std::vector<int64_t>* last = nullptr;
    std::array<int64_t, 6> last_key = {0, 0, 0, 0, 0, 0};
    for (size_t r = 0; r < n_rows; ++r) {
      const ItemDev& it = B.rows[r];
      if (!is_pair_family(it.family) || it.m_hi <= it.m_lo) continue;
      std::array<int64_t, 6> key = {it.tgt_dim, it.src_dim, it.m_lo, it.m_hi, it.wa_off, it.wb_off};
      // consecutive rows mostly share their streams: one map lookup per run of equal keys
      if (last == nullptr || key != last_key) {
        auto f = groups.find(key);
        if (f == groups.end()) {
          f = groups.emplace(key, found.size()).first;
          found.emplace_back();
          weighted.push_back(key[4] >= 0 || key[5] >= 0);
        }
        last = &found[f->second];   // (found may have grown: taken after the emplace)
        last_key = key;
      }
      last->push_back((int64_t)r);
    }
    // unweighted streams first: the two kinds are launched separately
    for (int pass = 0; pass < 2; ++pass)
      for (size_t g = 0; g < found.size(); ++g)
        if (weighted[g] == (pass == 1)) {
          B.q_group_stream.push_back(build_stream(ctx, B, found[g]));
          B.q_groups.emplace_back(std::move(found[g]));
        }
    // order the streams by category (stable) and renumber the groups' stream index
    auto category = [](const QStreamDev& S) {
      return (S.wa_off >= 0 || S.wb_off >= 0) ? 2 : (S.shape != LQ_SHAPE_BIG ? 0 : 1);
    };
    std::vector<int> perm(B.qstreams.size()), new_ix(B.qstreams.size());
    for (size_t x = 0; x < perm.size(); ++x) perm[x] = (int)x;
    auto arrival = [](const QStreamDev& S) { return std::max(S.tgt_dim, S.src_dim); };
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) {
      const int ca = category(B.qstreams[a]), cb = category(B.qstreams[b]);
      return ca != cb ? ca < cb : arrival(B.qstreams[a]) < arrival(B.qstreams[b]);
    });
    std::vector<QStreamDev> sorted(B.qstreams.size());
    for (size_t x = 0; x < perm.size(); ++x) {
      sorted[x] = B.qstreams[perm[x]];
      new_ix[perm[x]] = (int)x;
    }
    B.qstreams.swap(sorted);
    for (auto& sx : B.q_group_stream)
      if (sx >= 0) sx = new_ix[(size_t)sx];
  }
  // ---- quadrature streams: work units of whole tiles, a few waves over the 148 SMs ----------
  // (per category: the streams of one launch share the size of their work units)
  auto category = [](const QStreamDev& S) {
    return (S.wa_off >= 0 || S.wb_off >= 0) ? 2 : (S.shape != LQ_SHAPE_BIG ? 0 : 1);
  };
  int64_t tot[3] = {0, 0, 0}, tpw_cat[3];
  for (auto& S : B.qstreams) tot[category(S)] += S.m_hi - S.m_lo;
  for (int c = 0; c < 3; ++c) {
    const bool warp_shape = c == 0 && lq_warp_shape_enabled();
    const int64_t tm = warp_shape ? LQW_TM : (c == 0 ? LqSmall::TM : LqBig::TM);
    // (small CTAs: ~32 work units per resident CTA measured best on the C5 step, 10.42 ms against
    // 11.04 ms with 1.5 - 4 per CTA and 10.51 ms with 83: gpurun_out / profiles r2_g_units)
    int64_t units = c == 0 ? 148 * LqSmall::CTAS * 32 : 148 * 2 * 6;
    // ... but at least ~12 tiles per work unit: a unit pays for its first window search, the zeroing and
    // the write-out of its moment table (one eighth of C5, one time-block shard of eight: 1.372 ms with
    // 9 472 units, 1.422 ms with 18 944, 1.598 ms with 1 184)
    if (c == 0) units = std::min<int64_t>(units, std::max<int64_t>(148 * LqSmall::CTAS, tot[c] / (12 * tm)));
    // one warp per unit: 2 368 resident warps, ~12 units each, at least ~48 tiles of 32 targets per unit
    if (warp_shape)
      units = std::min<int64_t>(148 * LQW_WARPS * 12, std::max<int64_t>(148 * LQW_WARPS, tot[c] / (48 * tm)));
    if (const char* e = std::getenv("ASLSD_LQ_UNITS")) units = std::max<int64_t>(1, std::atoll(e));   // tuning
    int64_t tpw = (tot[c] + units - 1) / units;
    tpw = std::max<int64_t>(tm, std::min<int64_t>(tpw, 64 * LqBig::TM));
    tpw_cat[c] = ((tpw + tm - 1) / tm) * tm;
  }
  int64_t qwu = 0, part = 0, wts = 0;
  int cat_seen = 0;
  B.q_cat_wu[0] = 0;
  for (auto& S : B.qstreams) {
    const int c = category(S);
    while (cat_seen < c) B.q_cat_wu[++cat_seen] = qwu;
    S.tpw = tpw_cat[c];
    S.n_chunks = (int32_t)((S.m_hi - S.m_lo + S.tpw - 1) / S.tpw);
    S.wu_first = qwu;
    S.part_first = part;
    S.w_first = wts;
    qwu += S.n_chunks;
    part += (int64_t)S.n_chunks * S.J * LQ_ND;
    wts += (int64_t)S.J * LQ_ND;
  }
  while (cat_seen < 3) B.q_cat_wu[++cat_seen] = qwu;
  B.q_arrival_wu.assign((size_t)d + 1, B.q_cat_wu[1]);
  for (auto it = B.qstreams.rbegin(); it != B.qstreams.rend(); ++it)
    if (category(*it) == 0)
      for (int a = 0; a <= std::max(it->tgt_dim, it->src_dim); ++a) B.q_arrival_wu[(size_t)a] = it->wu_first;
  B.n_wu_q = qwu;
  B.n_q_partials = part;
  B.n_q_weights = wts;
  if (qwu >= (int64_t)2147483647) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: too many work units");
  return ASLSD_OK;
}

static int build_work_rows(aslsd_ctx* ctx, BatchPlan& B) {
  if (!B.rows_pending) return ASLSD_OK;
  const size_t n_rows = B.rows.size();
  std::vector<bool> fast(n_rows, false);
  std::vector<ItemDev>* qfam = B.scratch_q;
  for (int f = 0; f < ITEM_N_FAMILIES; ++f) qfam[f].clear();
  // rows of the streams: interior through the quadrature, targets near T_f through lag_boundary_kernel
  for (size_t g = 0; g < B.q_groups.size(); ++g) {
    const int sx = B.q_group_stream[g];
    if (sx < 0) continue;
    const QStreamDev& S = B.qstreams[(size_t)sx];
    for (int64_t r : B.q_groups[g]) {
      ItemDev it = B.rows[(size_t)r];
      it.dst_row = r;
      it.wu_first = sx;
      it.m_hi = S.m_hi;
      qfam[it.family].push_back(it);
      fast[(size_t)r] = true;
      const int64_t j_lo = std::max<int64_t>(0, (int64_t)std::floor(it.s_lo / S.h) - 1);
      const int64_t j_hi = std::min<int64_t>(S.J - 1, (int64_t)std::floor(it.s_hi / S.h) + 1);
      B.n_q_evals += std::max<int64_t>(0, j_hi - j_lo + 1) * LQ_ND;
    }
  }
  for (int f = 0; f < ITEM_N_FAMILIES; ++f) {
    B.qfam_first[f] = (int)B.qrows.size();
    B.qfam_count[f] = (int)qfam[f].size();
    B.bfam_first[f] = (int)B.bgroups.size();
    // rows of one stream are contiguous inside a family: one boundary group per (family, stream)
    // whose targets reach beyond the interior
    for (size_t x = 0; x < qfam[f].size();) {
      size_t y = x;
      while (y < qfam[f].size() && qfam[f][y].wu_first == qfam[f][x].wu_first) ++y;
      const QStreamDev& S = B.qstreams[(size_t)qfam[f][x].wu_first];
      const int64_t m_end = B.hi[S.tgt_dim];
      if (S.m_hi < m_end) {
        // a few CTAs per stream: each takes a slice of the rows (one warp per row at a time)
        const size_t rows_per_group = 4 * (LB_NT / 32);
        for (size_t z = x; z < y; z += rows_per_group) {
          BoundaryGroupDev G;
          std::memset(&G, 0, sizeof(G));
          G.stream = (int32_t)qfam[f][x].wu_first;
          G.row_first = (int32_t)(B.qrows.size() + z);
          G.row_count = (int32_t)std::min(rows_per_group, y - z);
          G.m_lo = S.m_hi;
          G.m_hi = m_end;
          B.bgroups.push_back(G);
        }
      }
      x = y;
    }
    B.bfam_count[f] = (int)B.bgroups.size() - B.bfam_first[f];
    B.qrows.insert(B.qrows.end(), qfam[f].begin(), qfam[f].end());
  }
  // ---- per-item kernels: rows outside the quadrature entirely, the others beyond the interior ----
  std::vector<ItemDev>* fam_items = B.scratch_f;
  for (int f = 0; f < ITEM_N_FAMILIES; ++f) fam_items[f].clear();
  for (size_t r = 0; r < n_rows; ++r) {
    ItemDev it = B.rows[r];
    it.dst_row = (int64_t)r;
    if (fast[r]) {
      continue;   // interior: lag quadrature; targets near T_f: lag_boundary_kernel
    } else if (!is_pair_family(it.family) && it.m_hi > it.m_lo && std::isfinite(it.t_int) && it.wa_off < 0) {
      // single-index rows are constant in the interior: n_int x closed form
      const int64_t m_b = interior_end(ctx, it.tgt_dim, it.m_lo, it.m_hi, it.t_int);
      it.n_int = m_b - it.m_lo;
      it.m_lo = m_b;
    }
    fam_items[it.family].push_back(it);
  }
  int base = 0;
  int64_t wu = 0;
  for (int f = 0; f < ITEM_N_FAMILIES; ++f) {
    B.fam_first[f] = base;
    B.fam_count[f] = (int)fam_items[f].size();
    B.fam_wu_first[f] = wu;
    const bool pair = is_pair_family(f);
    int64_t tot_targets = 0;
    for (auto& it : fam_items[f]) tot_targets += std::max<int64_t>(0, it.m_hi - it.m_lo);
    // enough work units for a few waves of 148 SMs x 2 CTAs, each spanning whole tiles
    int64_t tpw = (tot_targets + 148 * 2 * 8 - 1) / (148 * 2 * 8);
    const int64_t gran = pair ? PAIR_TM : 4 * PAIR_NT;
    tpw = std::max<int64_t>(gran, std::min<int64_t>(tpw, 16 * gran));
    tpw = ((tpw + gran - 1) / gran) * gran;
    for (auto& it : fam_items[f]) {
      it.tpw = tpw;
      it.n_chunks = it.m_hi > it.m_lo ? (int)((it.m_hi - it.m_lo + tpw - 1) / tpw) : 0;
      if (it.n_chunks == 0 && it.n_int > 0) it.n_chunks = 1;   // a CTA to add the interior constant
      it.wu_first = wu;
      wu += it.n_chunks;
      B.items.push_back(it);
    }
    B.fam_wu_count[f] = wu - B.fam_wu_first[f];
    base += B.fam_count[f];
  }
  B.n_wu = wu;
  B.rows_pending = false;
  if (B.n_wu >= (int64_t)2147483647) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: too many work units");
  return ASLSD_OK;
}

static int build_work(aslsd_ctx* ctx, const int64_t* m_lo, const int64_t* m_hi, BatchPlan& B, bool defer_rows) {
  int rc;
  if ((rc = build_work_streams(ctx, m_lo, m_hi, B))) return rc;
  return defer_rows ? ASLSD_OK : build_work_rows(ctx, B);
}

// The last batch plan is kept.  Its rows are a pure function of (model, flags, sub-problems,
// non-omega parameters); its work part also of (events, target block, options).  Each part is
// rebuilt only when its inputs change, and the omega-dependent variables are refreshed on
// every call.  No result is cached - every call launches all of its kernels.
static void free_batch(void* p) { delete (PlanCache*)p; }

static int get_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs, int flags,
                     const int64_t* m_lo, const int64_t* m_hi, bool with_work, BatchPlan** out) {
  {
    // a pending order check of the upload (page-locked inputs): take the verdict if it is there,
    // else go on - the kernels that rely on sorted times return at once on unsorted ones
    const int rc = events_verdict(ctx, false);
    if (rc) return rc;
  }
  std::vector<double> key_rows, key_work;
  key_rows.push_back((double)ctx->model_epoch);
  key_rows.push_back((double)n_k);
  key_rows.push_back((double)flags);
  for (int q = 0; q < n_k; ++q) {
    key_rows.push_back((double)ks[q]);
    plan_shape_key(ctx, ks[q], xs[q], key_rows);
  }
  key_work.push_back((double)ctx->events_epoch);
  key_work.push_back(ctx->use_quadrature ? 1. : 0.);
  for (int i = 0; i < ctx->d; ++i) {
    key_work.push_back(m_lo ? (double)m_lo[i] : -1.);
    key_work.push_back(m_hi ? (double)m_hi[i] : -1.);
  }
  if (!ctx->plan_cache) {
    ctx->plan_cache = new PlanCache();
    ctx->plan_cache_free = free_batch;
  }
  PlanCache& cache = *(PlanCache*)ctx->plan_cache;
  // plans of an earlier model epoch can never match again (set_model, set_option, weighted calls)
  for (size_t x = 0; x < cache.plans.size();)
    if (cache.plans[x]->key_rows[0] != (double)ctx->model_epoch) {
      delete cache.plans[x];
      cache.plans.erase(cache.plans.begin() + x);
    } else {
      ++x;
    }
  BatchPlan* C = nullptr;
  for (size_t x = 0; x < cache.plans.size(); ++x)
    if (cache.plans[x]->key_rows == key_rows) {
      C = cache.plans[x];
      cache.plans.erase(cache.plans.begin() + x);
      cache.plans.insert(cache.plans.begin(), C);
      break;
    }
  int rc;
  if (C) {
    for (int q = 0; q < n_k; ++q)
      if ((rc = plan_vars(ctx, ks[q], xs[q], C->plans[q]))) return rc;
  } else {
    BatchPlan* B = new BatchPlan();
    B->uid = ++ctx->plan_uid;
    if ((rc = build_rows(ctx, n_k, ks, xs, flags, *B))) { delete B; return rc; }
    B->key_rows.swap(key_rows);
    cache.plans.insert(cache.plans.begin(), C = B);
    while (cache.plans.size() > PLAN_CACHE_MAX) {
      delete cache.plans.back();
      cache.plans.pop_back();
    }
  }
  if (with_work && !(C->has_work && C->key_work == key_work)) {
    C->has_work = false;
    const auto t0 = std::chrono::steady_clock::now();
    if ((rc = build_work(ctx, m_lo, m_hi, *C, true))) return rc;   // the rows: in run_batch, behind the first launches
    if (std::getenv("ASLSD_TRACE"))
      std::fprintf(stderr, "[aslsd] build_work %.3f ms (%zu rows, %zu streams)\n",
                   1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(),
                   C->rows.size(), C->qstreams.size());
    C->key_work.swap(key_work);
    C->has_work = true;
  }
  *out = C;
  return ASLSD_OK;
}

// Do the node weights in d_q_weights (left by another plan) serve this plan?  They are a function
// of the event streams, the target block and the bin grid of every stream: same events, same
// streams in the same order with the same geometry, no per-event weights.
static bool weights_match(const aslsd_ctx* ctx, const BatchPlan& B) {
  if (ctx->w_events_epoch != ctx->events_epoch) return false;
  if (ctx->w_streams.size() != B.qstreams.size() * sizeof(QStreamDev)) return false;
  const QStreamDev* W = (const QStreamDev*)ctx->w_streams.data();
  for (size_t x = 0; x < B.qstreams.size(); ++x) {
    const QStreamDev& a = W[x];
    const QStreamDev& b = B.qstreams[x];
    if (a.tgt_dim != b.tgt_dim || a.src_dim != b.src_dim || a.same != b.same || a.J != b.J || a.m_lo != b.m_lo ||
        a.m_hi != b.m_hi || a.w_first != b.w_first || a.s_max != b.s_max || a.h != b.h || a.wa_off >= 0 ||
        a.wb_off >= 0 || b.wa_off >= 0 || b.wb_off >= 0)
      return false;
  }
  return true;
}

template <int FAMILY>
static void launch_pair(aslsd_ctx* ctx, const BatchPlan& B, const ItemDev* d_items, cudaStream_t st) {
  if (B.fam_wu_count[FAMILY] == 0) return;
  pair_kernel<FAMILY><<<(unsigned)B.fam_wu_count[FAMILY], PAIR_NT, 0, st>>>(
      (const DimDev*)ctx->d_dims, d_items + B.fam_first[FAMILY], B.fam_count[FAMILY], ctx->T_f,
      ctx->d_partials, ctx->d_counts, B.fam_wu_first[FAMILY], ctx->d_weights);
  ctx->last_launches++;
}
template <int FAMILY>
static void launch_single(aslsd_ctx* ctx, const BatchPlan& B, const ItemDev* d_items, cudaStream_t st) {
  if (B.fam_wu_count[FAMILY] == 0) return;
  single_kernel<FAMILY><<<(unsigned)B.fam_wu_count[FAMILY], PAIR_NT, 0, st>>>(
      (const DimDev*)ctx->d_dims, d_items + B.fam_first[FAMILY], B.fam_count[FAMILY], ctx->T_f,
      ctx->d_partials, ctx->d_counts, B.fam_wu_first[FAMILY], ctx->d_weights);
  ctx->last_launches++;
}

static int upload_quadrature_tables(aslsd_ctx* ctx) {
  if (ctx->lq_tables_set) return ASLSD_OK;
  double x[LQ_ND], ct[LQ_ND * LQ_ND];
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < LQ_ND; ++i) {
    x[i] = std::cos(pi * (i + 0.5) / LQ_ND);
    for (int d = 0; d < LQ_ND; ++d)
      ct[i * LQ_ND + d] = (d == 0 ? 1. : 2.) / LQ_ND * std::cos(pi * d * (i + 0.5) / LQ_ND);
  }
  CUDA_TRY(ctx, cudaMemcpyToSymbol(LQ_X, x, sizeof(x)));
  CUDA_TRY(ctx, cudaMemcpyToSymbol(LQ_CT, ct, sizeof(ct)));
  ctx->lq_tables_set = true;
  return ASLSD_OK;
}

template <int FAMILY>
static void launch_quad(aslsd_ctx* ctx, const BatchPlan& B, const QStreamDev* d_streams, const ItemDev* d_qrows,
                        double* raw_device, cudaStream_t st) {
  if (B.qfam_count[FAMILY] == 0) return;
  lag_quad_kernel<FAMILY><<<(unsigned)B.qfam_count[FAMILY], LQ_QT, 0, st>>>(
      d_streams, d_qrows + B.qfam_first[FAMILY], ctx->d_q_weights, raw_device);
  ctx->last_launches++;
}

template <int FAMILY>
static void launch_boundary(aslsd_ctx* ctx, const BatchPlan& B, const QStreamDev* d_streams,
                            const BoundaryGroupDev* d_groups, const ItemDev* d_qrows, double* raw_device,
                            unsigned long long* d_cnt, cudaStream_t st) {
  if (B.bfam_count[FAMILY] == 0) return;
  lag_boundary_kernel<FAMILY><<<(unsigned)B.bfam_count[FAMILY], LB_NT, 0, st>>>(
      (const DimDev*)ctx->d_dims, d_streams, d_groups + B.bfam_first[FAMILY], d_qrows, ctx->T_f, raw_device,
      d_cnt + B.bfam_first[FAMILY], ctx->d_weights);
  ctx->last_launches++;
}

// Launch everything; leaves raw[row][comp] in raw_device (zeroed here first).
static int run_batch_impl(aslsd_ctx* ctx, BatchPlan& B, double* raw_device);
static int run_batch(aslsd_ctx* ctx, BatchPlan& B, double* raw_device) {
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = run_batch_impl(ctx, B, raw_device);
  if (std::getenv("ASLSD_TRACE"))
    std::fprintf(stderr, "[aslsd] run_batch host side %.3f ms\n",
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  return rc;
}
// The moment kernels of every category and the weights kernel, on the context's stream.
// `st`: the compute stream.  With event uploads still in flight on the copy stream (`staged`, st then is
// the raw stream, not yet ordered behind them) the small-shape launch is cut by the later event type of
// its streams: each piece waits for that type's upload only and runs while the following ones arrive;
// the stream is ordered behind all uploads before anything else is launched.
static int launch_moments(aslsd_ctx* ctx, BatchPlan& B, const QStreamDev* d_streams, cudaStream_t st, bool staged) {
  if (ctx->q_smem_set == 0) {
    // the attribute belongs to the function, not to this context: every context asks for the
    // same ceiling, so contexts with different bin counts never lower each other's limit
    CUDA_TRY(ctx, cudaFuncSetAttribute(lag_moment_kernel<false, LqSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)LqSmall::SMEM_BUDGET));
    CUDA_TRY(ctx, cudaFuncSetAttribute(lag_moment_kernel<false, LqBig>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)LqBig::SMEM_BUDGET));
    CUDA_TRY(ctx, cudaFuncSetAttribute(lag_moment_kernel<true, LqBig>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)LqBig::SMEM_BUDGET));
    // (static shared memory, 16 one-warp CTAs per SM: the largest shared-memory carve-out)
    CUDA_TRY(ctx, cudaFuncSetAttribute(lag_moment_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
    ctx->q_smem_set = LqBig::SMEM_BUDGET;
  }
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_q0, st));
  unsigned long long* d_qcnt = ctx->d_qcounts;
  const int n_streams = (int)B.qstreams.size();
  // category 0: unweighted streams of at most 64 bins, one warp or one 128-thread CTA per work unit
  const bool warp_shape = !B.qstreams.empty() && B.qstreams[0].shape == LQ_SHAPE_WARP;
  auto launch_small = [&](int64_t w0, int64_t w1, cudaStream_t on) {
    if (warp_shape)
      lag_moment_warp_kernel<<<(unsigned)(w1 - w0), 32, 0, on>>>((const DimDev*)ctx->d_dims, d_streams, n_streams,
                                                                 ctx->d_q_partials, d_qcnt, w0);
    else
      lag_moment_kernel<false, LqSmall><<<(unsigned)(w1 - w0), LqSmall::NT, B.q_cat_smem[0], on>>>(
          (const DimDev*)ctx->d_dims, d_streams, n_streams, ctx->d_q_partials, d_qcnt, ctx->d_weights, w0);
    ctx->last_launches++;
  };
  if (B.q_cat_wu[1] > B.q_cat_wu[0]) {
    if (!staged) {
      launch_small(B.q_cat_wu[0], B.q_cat_wu[1], st);
    } else {
      // pieces of at least 1/16 of the launch, alternating between the compute stream and a second
      // lane: the last wave of a piece (a work unit runs ~0.3 ms at C5) is filled by the next piece
      const int64_t total = B.q_cat_wu[1] - B.q_cat_wu[0];
      cudaStream_t lanes[2] = {st, ctx->aux_stream};
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_aux0, st));                 // the stream table is on its way on st
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_aux0, 0));
      int piece = 0;
      int64_t w0 = B.q_cat_wu[0];
      for (int a = 0; a < ctx->n_upload_types; ++a) {
        const int64_t w1 = B.q_arrival_wu[(size_t)a + 1];
        const bool last = a + 1 == ctx->n_upload_types;
        if (w1 <= w0 || (!last && w1 - w0 < total / 16)) continue;    // too small: joins the next type's piece
        cudaStream_t lane = lanes[piece++ & 1];
        CUDA_TRY(ctx, cudaStreamWaitEvent(lane, ctx->ev_type[(size_t)a], 0));
        launch_small(w0, w1, lane);
        w0 = w1;
      }
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_aux1, ctx->aux_stream));
      CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_aux1, 0));
    }
  }
  if (staged) {
    aslsd_join_uploads(ctx);   // everything below may read any event type
    st = ctx->stream;
  }
  if (B.q_cat_wu[2] > B.q_cat_wu[1]) {
    lag_moment_kernel<false, LqBig><<<(unsigned)(B.q_cat_wu[2] - B.q_cat_wu[1]), LqBig::NT, B.q_cat_smem[1], st>>>(
        (const DimDev*)ctx->d_dims, d_streams, n_streams, ctx->d_q_partials, d_qcnt, ctx->d_weights, B.q_cat_wu[1]);
    ctx->last_launches++;
  }
  if (B.q_cat_wu[3] > B.q_cat_wu[2]) {
    lag_moment_kernel<true, LqBig><<<(unsigned)(B.q_cat_wu[3] - B.q_cat_wu[2]), LqBig::NT, B.q_cat_smem[2], st>>>(
        (const DimDev*)ctx->d_dims, d_streams, n_streams, ctx->d_q_partials, d_qcnt, ctx->d_weights, B.q_cat_wu[2]);
    ctx->last_launches++;
  }
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_q1, st));
  ctx->moment_timing_pending = true;
  const dim3 wgrid((unsigned)((B.q_jmax + LQ_WB - 1) / LQ_WB), (unsigned)B.qstreams.size());
  ctx->w_streams.clear();   // being rewritten
  lag_weights_kernel<<<wgrid, LQ_WB * LQ_ND, 0, st>>>(d_streams, ctx->d_q_partials, ctx->d_q_weights);
  ctx->last_launches++;
  if (ctx->cache_weights) {   // record what the buffer holds: later calls (of any plan) may reuse it
    ctx->w_streams.assign((const unsigned char*)B.qstreams.data(),
                          (const unsigned char*)(B.qstreams.data() + B.qstreams.size()));
    ctx->w_events_epoch = ctx->events_epoch;
  }
  return ASLSD_OK;
}

static int run_batch_impl(aslsd_ctx* ctx, BatchPlan& B, double* raw_device) {
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t n_rows = B.rows.size();
  int rc;
  ctx->last_launches = 0;
  ctx->moment_timing_pending = false;
  ctx->last_moment_ms = 0.;
  const bool quad = !B.qstreams.empty();
  if (ctx->tables_owner != (const void*)&B || ctx->tables_owner_uid != B.uid) B.device_resident = false;
  // device table block [streams | rows | boundary groups], page-locked staging block
  // [streams | rows | boundary groups | items]: the streams come first in both, their size is known
  // before the rows are planned
  const size_t o_s = 0;
  const size_t o_r = o_s + ((B.qstreams.size() * sizeof(QStreamDev) + 15) & ~(size_t)15);
  bool streams_uploaded = false, moments_done = false;
  // (raw: recording the start of the call must not order the stream behind uploads in flight)
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream.raw()));
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream.raw()));
  if (quad) {
    if ((rc = upload_quadrature_tables(ctx))) return rc;
    if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_q_partials, &ctx->q_partials_cap,
                                   (size_t)B.n_q_partials * sizeof(double)))) return rc;
    const double* before = ctx->d_q_weights;
    if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_q_weights, &ctx->q_weights_cap,
                                   (size_t)B.n_q_weights * sizeof(double)))) return rc;
    if (ctx->d_q_weights != before) ctx->w_streams.clear();
    if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_qcounts, &ctx->qcounts_cap,
                                   ((size_t)B.n_wu_q + 1) * sizeof(unsigned long long)))) return rc;
  }
  // The node weights depend on the event streams and on the bin grids only (not on omega, not
  // on the evaluation point as long as the windows stay): with the weight cache on, a repeated
  // evaluation on the same plan - the iterations of a fit with fixed (beta, delta) - skips the
  // pass over the event pairs.
  // (the record of what d_q_weights holds decides: several plans share the buffer)
  const bool reuse_weights = quad && ctx->cache_weights && weights_match(ctx, B);
  ctx->last_weights_reused = reuse_weights;
  if (B.rows_pending && quad && !B.device_resident) {
    // a fresh work part: streams to the device and the moment kernels on their way, then the rows are
    // planned while they run.  The blocks are sized for the largest row part this plan can have.
    const size_t n_groups_max = n_rows + (size_t)ITEM_N_FAMILIES * B.qstreams.size();
    const size_t q_bytes_max = o_r + ((n_rows * sizeof(ItemDev) + 15) & ~(size_t)15) +
                               n_groups_max * sizeof(BoundaryGroupDev);
    const size_t need = q_bytes_max + ((n_rows * sizeof(ItemDev) + 15) & ~(size_t)15);
    if (ctx->tables_in_flight) {
      CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_tables));
      ctx->tables_in_flight = false;
    }
    if (need > ctx->tables_cap) {
      if (ctx->h_tables) CUDA_TRY(ctx, cudaFreeHost(ctx->h_tables));
      ctx->h_tables = nullptr;
      ctx->tables_cap = 0;
      CUDA_TRY(ctx, cudaMallocHost(&ctx->h_tables, need + need / 4));
      ctx->tables_cap = need + need / 4;
    }
    if ((rc = ensure_device_buffer(ctx, &ctx->d_qtab, &ctx->qtab_cap, q_bytes_max))) return rc;
    char* hp = (char*)ctx->h_tables;
    std::memcpy(hp + o_s, B.qstreams.data(), B.qstreams.size() * sizeof(QStreamDev));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_qtab, hp, o_r, cudaMemcpyHostToDevice, ctx->stream.raw()));
    streams_uploaded = true;
    if (!reuse_weights) {
      // (worth it when the transfer is long against the planning that already overlaps it: >= 24 MB)
      const bool staged = ctx->uploads_pending && ctx->n_upload_types == ctx->d && ctx->upload_bytes >= ((size_t)24 << 20);
      if ((rc = launch_moments(ctx, B, (const QStreamDev*)((const char*)ctx->d_qtab + o_s),
                               staged ? ctx->stream.raw() : (cudaStream_t)ctx->stream, staged)))
        return rc;
      moments_done = true;
    }
  }
  if ((rc = build_work_rows(ctx, B))) return rc;
  const size_t n_items = B.items.size();
  ctx->last_n_wu = B.n_wu;
  ctx->last_n_items = (int)n_items;
  ctx->last_n_wu_q = B.n_wu_q;
  for (int f = 0; f < ITEM_N_FAMILIES; ++f) { ctx->last_fam_first[f] = B.fam_first[f]; ctx->last_fam_count[f] = B.fam_count[f]; }
  ctx->last_work[3] = B.n_q_evals;
  ctx->last_work[4] = (int64_t)B.qstreams.size();
  ctx->last_work[5] = B.n_q_weights / LQ_ND;
  ctx->last_work[6] = (int64_t)n_rows;
  ctx->last_work[7] = B.n_items_total;
  if ((rc = ensure_device_buffer(ctx, &ctx->d_items, &ctx->items_cap, std::max<size_t>(1, n_items) * sizeof(ItemDev))))
    return rc;
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_partials, &ctx->partials_cap,
                                 std::max<size_t>(1, (size_t)B.n_wu) * ASLSD_MAX_COMP * sizeof(double)))) return rc;
  // counters: [per-item work units | items | boundary groups]; the moment kernels have their own block
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_counts, &ctx->counts_cap,
                                 ((size_t)B.n_wu + n_items + B.bgroups.size() + 1) * sizeof(unsigned long long))))
    return rc;
  const size_t o_g = o_r + ((B.qrows.size() * sizeof(ItemDev) + 15) & ~(size_t)15);
  const size_t q_bytes = o_g + B.bgroups.size() * sizeof(BoundaryGroupDev);
  // Two lanes.  The moment kernels and the node weights need the event pairs and nothing else; the per-item
  // kernels and the targets near T_f need the rows and nothing of the moments: when the moments run, those
  // take the second lane (`side`) - they fill the last, partly empty wave of the moment kernels instead of
  // following it - and the quadrature rows, which need both, wait for it.  Every row is written in a fixed
  // order whatever the lanes do: zero, per-item partials (finalize), targets near T_f, quadrature.
  const bool fork = quad && !reuse_weights;
  cudaStream_t main_st = ctx->stream;
  cudaStream_t side = fork ? ctx->aux_stream : main_st;
  // (ev_fork: what the side lane starts behind - recorded at the top of this call when the moment kernels
  // are already out, so that it waits for the previous call's readers of the tables and of the raw rows
  // but not for them; otherwise recorded below, behind the upload of the tables)
  if (!B.device_resident && (n_items || quad)) {
    // all plan tables go through one page-locked block of their own and asynchronous transfers:
    // no wait for the stream here (an event upload may still be in flight on it)
    const size_t i_bytes = (n_items * sizeof(ItemDev) + 15) & ~(size_t)15;
    const size_t need = i_bytes + (quad ? q_bytes : 0);
    if (!streams_uploaded) {
      if (ctx->tables_in_flight) {
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_tables));
        ctx->tables_in_flight = false;
      }
      if (need > ctx->tables_cap) {
        if (ctx->h_tables) CUDA_TRY(ctx, cudaFreeHost(ctx->h_tables));
        ctx->h_tables = nullptr;
        ctx->tables_cap = 0;
        CUDA_TRY(ctx, cudaMallocHost(&ctx->h_tables, need + need / 4));
        ctx->tables_cap = need + need / 4;
      }
    }
    // with the moment kernels already on the compute stream the rest of the tables travels on the side
    // lane, which reads them
    cudaStream_t up = moments_done ? side : main_st;
    if (moments_done) CUDA_TRY(ctx, cudaStreamWaitEvent(side, ctx->ev_fork, 0));
    char* hp = (char*)ctx->h_tables;
    char* hi = hp + (quad ? q_bytes : 0);   // items behind the quadrature tables
    if (quad) {
      if (!streams_uploaded) {
        if ((rc = ensure_device_buffer(ctx, &ctx->d_qtab, &ctx->qtab_cap, q_bytes))) return rc;
        std::memcpy(hp + o_s, B.qstreams.data(), B.qstreams.size() * sizeof(QStreamDev));
      }
      std::memcpy(hp + o_r, B.qrows.data(), B.qrows.size() * sizeof(ItemDev));
      std::memcpy(hp + o_g, B.bgroups.data(), B.bgroups.size() * sizeof(BoundaryGroupDev));
      const size_t from = streams_uploaded ? o_r : 0;
      if (q_bytes > from)
        CUDA_TRY(ctx, cudaMemcpyAsync((char*)ctx->d_qtab + from, hp + from, q_bytes - from, cudaMemcpyHostToDevice, up));
    }
    if (n_items) {
      std::memcpy(hi, B.items.data(), n_items * sizeof(ItemDev));
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_items, hi, n_items * sizeof(ItemDev), cudaMemcpyHostToDevice, up));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_tables, up));
    ctx->tables_in_flight = true;
  }
  B.device_resident = true;
  ctx->tables_owner = &B;
  ctx->tables_owner_uid = B.uid;
  const ItemDev* d_items = (const ItemDev*)ctx->d_items;
  const QStreamDev* d_streams = (const QStreamDev*)((const char*)ctx->d_qtab + o_s);
  const ItemDev* d_qrows = (const ItemDev*)((const char*)ctx->d_qtab + o_r);
  const BoundaryGroupDev* d_groups = (const BoundaryGroupDev*)((const char*)ctx->d_qtab + o_g);
  if (fork && !moments_done) {
    // tables uploaded on the compute stream: the side lane starts behind them, the moment kernels follow
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, main_st));
    if ((rc = launch_moments(ctx, B, d_streams, main_st, false))) return rc;
  }
  if (fork) CUDA_TRY(ctx, cudaStreamWaitEvent(side, ctx->ev_fork, 0));
  CUDA_TRY(ctx, cudaMemsetAsync(raw_device, 0, std::max<size_t>(1, n_rows) * ASLSD_MAX_COMP * sizeof(double), side));
  launch_pair<ITEM_UPS_GAUSS>(ctx, B, d_items, side);
  launch_pair<ITEM_UPS_EXP>(ctx, B, d_items, side);
  launch_pair<ITEM_UPS_EG>(ctx, B, d_items, side);
  launch_pair<ITEM_UPS_GE>(ctx, B, d_items, side);
  launch_pair<ITEM_PHI_GAUSS>(ctx, B, d_items, side);
  launch_pair<ITEM_PHI_EXP>(ctx, B, d_items, side);
  launch_single<ITEM_UPS0_GAUSS>(ctx, B, d_items, side);
  launch_single<ITEM_UPS0_EXP>(ctx, B, d_items, side);
  launch_single<ITEM_UPS0_EG>(ctx, B, d_items, side);
  launch_single<ITEM_UPS0_GE>(ctx, B, d_items, side);
  launch_single<ITEM_PSI_GAUSS>(ctx, B, d_items, side);
  launch_single<ITEM_PSI_EXP>(ctx, B, d_items, side);
  if (n_items) {
    const int n_fin = (int)n_items * (ASLSD_MAX_COMP + 1);
    finalize_kernel<<<(n_fin + 127) / 128, 128, 0, side>>>(d_items, (int)n_items, ctx->d_partials, ctx->d_counts,
                                                          raw_device, ctx->d_counts + B.n_wu);
    ctx->last_launches++;
  }
  if (quad) {
    unsigned long long* d_bcnt = ctx->d_counts + B.n_wu + n_items;
    launch_boundary<ITEM_UPS_GAUSS>(ctx, B, d_streams, d_groups, d_qrows, raw_device, d_bcnt, side);
    launch_boundary<ITEM_UPS_EXP>(ctx, B, d_streams, d_groups, d_qrows, raw_device, d_bcnt, side);
    launch_boundary<ITEM_UPS_EG>(ctx, B, d_streams, d_groups, d_qrows, raw_device, d_bcnt, side);
    launch_boundary<ITEM_UPS_GE>(ctx, B, d_streams, d_groups, d_qrows, raw_device, d_bcnt, side);
    launch_boundary<ITEM_PHI_GAUSS>(ctx, B, d_streams, d_groups, d_qrows, raw_device, d_bcnt, side);
    launch_boundary<ITEM_PHI_EXP>(ctx, B, d_streams, d_groups, d_qrows, raw_device, d_bcnt, side);
    if (fork) {
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, side));
      CUDA_TRY(ctx, cudaStreamWaitEvent(main_st, ctx->ev_join, 0));
    }
    launch_quad<ITEM_UPS_GAUSS>(ctx, B, d_streams, d_qrows, raw_device, main_st);
    launch_quad<ITEM_UPS_EXP>(ctx, B, d_streams, d_qrows, raw_device, main_st);
    launch_quad<ITEM_UPS_EG>(ctx, B, d_streams, d_qrows, raw_device, main_st);
    launch_quad<ITEM_UPS_GE>(ctx, B, d_streams, d_qrows, raw_device, main_st);
    launch_quad<ITEM_PHI_GAUSS>(ctx, B, d_streams, d_qrows, raw_device, main_st);
    launch_quad<ITEM_PHI_EXP>(ctx, B, d_streams, d_qrows, raw_device, main_st);
  }
  ctx->last_n_bgroups = (int64_t)B.bgroups.size();
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timing_pending = true;
  CUDA_TRY(ctx, cudaGetLastError());
  return ASLSD_OK;
}

// lo / hi: target block per dimension
static int count_pair_terms(aslsd_ctx* ctx, const SumPlan& P, const std::vector<int64_t>& lo,
                            const std::vector<int64_t>& hi, int64_t* n_pair_terms) {
  const size_t nt = P.terms.size();
  unsigned long long* d_cnt = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&d_cnt, nt * sizeof(unsigned long long)));
  cudaMemsetAsync(d_cnt, 0, nt * sizeof(unsigned long long), ctx->stream);
  for (size_t t = 0; t < nt; ++t) {
    const TermPlan& T = P.terms[t];
    if (T.kind < 2 || P.term_s_hi[t] < 0.) continue;
    int tgt = T.kind == 2 ? P.k : T.i;
    int src = T.kind == 2 ? T.i : T.j;
    const int64_t m_lo = lo[tgt], m_hi = hi[tgt];
    if (m_hi <= m_lo) continue;
    int64_t blocks = std::min<int64_t>((m_hi - m_lo + 255) / 256, 148 * 8);
    count_pairs_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const DimDev*)ctx->d_dims, tgt, src,
                                                                  tgt == src, m_lo, m_hi, P.term_s_lo[t],
                                                                  P.term_s_hi[t], d_cnt + t);
  }
  std::vector<unsigned long long> h(nt);
  cudaError_t e = cudaMemcpyAsync(h.data(), d_cnt, nt * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                  ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_cnt);
  CUDA_TRY(ctx, e);
  for (size_t t = 0; t < nt; ++t) n_pair_terms[t] = (int64_t)h[t];
  return ASLSD_OK;
}

// ---- assembly: raw rows -> packed per-term sums of one sub-problem -------------------------
static void assemble(const aslsd_ctx* ctx, const SumPlan& P, const std::vector<int64_t>& item_row,
                     const double* x_k, const double* raw, double* out) {
  for (int64_t c = 0; c < P.out_len; ++c) out[c] = 0.;
  for (size_t t = 0; t < P.terms.size(); ++t) {
    double* o = out + P.terms[t].out_off;
    emit_term(ctx, P, t, x_k, [&](int c, int item, int comp, double coef, double outer) {
      o[c] += outer * (coef * raw[(size_t)item_row[item] * ASLSD_MAX_COMP + comp]);
    });
  }
}

// The sub-problems of a batch are assembled independently: one host thread each once there is
// enough of them to pay for the threads (C5: 10 x ~2.7e5 coefficient products, 0.9 ms serial).
static void assemble_all(const aslsd_ctx* ctx, const BatchPlan& B, int n_k, const double* const* xs,
                         const double* raw, double* const* outs) {
  int64_t slots = 0;
  for (int q = 0; q < n_k; ++q) slots += (int64_t)B.plans[q].slot_item.size();
  if (n_k < 2 || slots < 2000) {
    for (int q = 0; q < n_k; ++q) assemble(ctx, B.plans[q], B.item_row[q], xs[q], raw, outs[q]);
    return;
  }
  // (4: several ranks of a box run this at the same time)
  const int n_thr = std::min<int>(n_k, std::max(1u, std::min(4u, std::thread::hardware_concurrency())));
  std::atomic<int> next(0);
  auto worker = [&]() {
    for (int q = next.fetch_add(1); q < n_k; q = next.fetch_add(1))
      assemble(ctx, B.plans[q], B.item_row[q], xs[q], raw, outs[q]);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_thr; ++t) pool.emplace_back(worker);
  worker();
  for (auto& th : pool) th.join();
}


// ---- assembly on the device -----------------------------------------------------------------
// The map is linear in the raw rows with coefficients that are products of at most two omegas
// (and mu_k) times a function of the non-omega variables: its structure is extracted once per row
// plan (emit_term_w<SymW>), the omegas travel as a small table per call.  One thread per packed
// sum walks its entries in emission order with the host's operations (no contraction into fused
// multiply-adds): bit for bit the packed sums of the host assembly.
#define ASM_NT 64
__device__ __forceinline__ double asm_term(const AsmEntry& E, const double* __restrict__ wtab,
                                           const double* __restrict__ raw) {
  const double coef = __ddiv_rn(__dmul_rn(__dmul_rn(wtab[E.a], wtab[E.b]), E.g), E.gd);
  return __dmul_rn(wtab[E.o], __dmul_rn(coef, raw[E.raw_ix]));
}
__global__ void __launch_bounds__(ASM_NT)
assemble_kernel(const int64_t* __restrict__ ptr, const AsmEntry* __restrict__ entries,
                const double* __restrict__ wtab, const double* __restrict__ raw, double* __restrict__ out,
                int64_t n_out) {
  const int64_t o = (int64_t)blockIdx.x * ASM_NT + threadIdx.x;
  if (o >= n_out) return;
  double acc = 0.;
  int64_t e = ptr[o];
  const int64_t e1 = ptr[o + 1];
  for (; e + 4 <= e1; e += 4) {   // the four terms are independent of the running sum
    const double v0 = asm_term(entries[e], wtab, raw), v1 = asm_term(entries[e + 1], wtab, raw);
    const double v2 = asm_term(entries[e + 2], wtab, raw), v3 = asm_term(entries[e + 3], wtab, raw);
    acc = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(acc, v0), v1), v2), v3);
  }
  for (; e < e1; ++e) acc = __dadd_rn(acc, asm_term(entries[e], wtab, raw));
  out[o] = acc;
}

// omega table of sub-problem q: [0] = 1, [1] = mu_k, then the omegas of K[k][i] read with its own
// structure (P.V) and with the structure of KU (P.VU)
static int asm_r_max(const aslsd_ctx* ctx) {
  size_t r = 1;
  for (auto& K : ctx->kernels) r = std::max(r, K.bases.size());
  return (int)r;
}
static void fill_wtab(const aslsd_ctx* ctx, const BatchPlan& B, int q, const double* x_k, double* w) {
  const int d = ctx->d, R = asm_r_max(ctx);
  const SumPlan& P = B.plans[q];
  w[0] = 1.;
  w[1] = x_k[0];
  for (int i = 0; i < d; ++i)
    for (int r = 0; r < R; ++r) {
      w[2 + i * R + r] = r < (int)P.V[i].size() ? P.V[i][r][0] : 0.;
      w[2 + d * R + i * R + r] = r < (int)P.VU[i].size() ? P.VU[i][r][0] : 0.;
    }
}

static int build_assembly_host(aslsd_ctx* ctx, BatchPlan& B, std::vector<int64_t>& count,
                               std::vector<AsmEntry>& sorted) {
  const int d = ctx->d, R = asm_r_max(ctx), n_k = (int)B.plans.size();
  B.w_stride = 2 + 2 * d * R;
  B.out_off.assign(n_k + 1, 0);
  for (int q = 0; q < n_k; ++q) B.out_off[q + 1] = B.out_off[q] + B.plans[q].out_len;
  B.out_total = B.out_off[n_k];
  if ((int64_t)B.rows.size() * ASLSD_MAX_COMP >= 2147483647LL || (int64_t)n_k * B.w_stride >= 2147483647LL)
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: assembly table too large");
  count.assign((size_t)B.out_total + 1, 0);
  struct Tmp { int64_t out; AsmEntry e; };
  std::vector<Tmp> tmp;
  for (int q = 0; q < n_k; ++q) {
    const SumPlan& P = B.plans[q];
    const int32_t base = q * B.w_stride;
    const std::vector<int64_t>& item_row = B.item_row[q];
    for (size_t t = 0; t < P.terms.size(); ++t) {
      const int64_t o0 = B.out_off[q] + P.terms[t].out_off;
      emit_term_w<SymW>(ctx, P, t, SymW(1, 1.),
                        [&](int i, int r) { return SymW(2 + i * R + r, 1.); },
                        [&](int i, int r) { return SymW(2 + d * R + i * R + r, 1.); },
                        [&](int c, int item, int comp, SymW coef, SymW outer) {
                          Tmp x;
                          x.out = o0 + c;
                          x.e.raw_ix = (int32_t)(item_row[item] * ASLSD_MAX_COMP + comp);
                          x.e.a = base + coef.a;
                          x.e.b = base + coef.b;
                          x.e.o = base + outer.a;
                          x.e.g = coef.g * outer.g;
                          x.e.gd = coef.gd * outer.gd;
                          tmp.push_back(x);
                          count[(size_t)x.out + 1]++;
                        });
    }
  }
  for (size_t o = 0; o < (size_t)B.out_total; ++o) count[o + 1] += count[o];
  sorted.resize(tmp.size());
  std::vector<int64_t> cur(count.begin(), count.end() - 1);
  for (auto& x : tmp) sorted[(size_t)cur[(size_t)x.out]++] = x.e;   // stable: emission order within a sum
  B.asm_entries = (int64_t)sorted.size();
  return ASLSD_OK;
}

static int build_assembly(aslsd_ctx* ctx, BatchPlan& B) {
  std::vector<int64_t> count;
  std::vector<AsmEntry> sorted;
  int rc;
  if ((rc = build_assembly_host(ctx, B, count, sorted))) return rc;
  if (B.d_asm) CUDA_TRY(ctx, cudaFree(B.d_asm));
  if (B.d_asm_ptr) CUDA_TRY(ctx, cudaFree(B.d_asm_ptr));
  B.d_asm = nullptr;
  B.d_asm_ptr = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&B.d_asm, std::max<size_t>(1, sorted.size()) * sizeof(AsmEntry)));
  CUDA_TRY(ctx, cudaMalloc((void**)&B.d_asm_ptr, count.size() * sizeof(int64_t)));
  CUDA_TRY(ctx, cudaMemcpyAsync(B.d_asm, sorted.data(), sorted.size() * sizeof(AsmEntry), cudaMemcpyHostToDevice,
                                ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(B.d_asm_ptr, count.data(), count.size() * sizeof(int64_t), cudaMemcpyHostToDevice,
                                ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // the sources are locals
  B.asm_built = true;
  return ASLSD_OK;
}

// raw_device (rows of this plan) -> packed_device (out_total doubles), on the context's stream
static int assemble_device(aslsd_ctx* ctx, BatchPlan& B, int n_k, const double* const* xs, const double* raw_device,
                           double* packed_device) {
  int rc;
  if (!B.asm_built)
    if ((rc = build_assembly(ctx, B))) return rc;
  const size_t n_w = (size_t)n_k * B.w_stride;
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_wtab, &ctx->wtab_cap, n_w * sizeof(double)))) return rc;
  // page-locked staging of the omega tables (two halves in turn: the previous call's transfer
  // may still be in flight when the next one is planned)
  if (n_w * 2 > ctx->h_wtab_cap) {
    if (ctx->h_wtab) {
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      CUDA_TRY(ctx, cudaFreeHost(ctx->h_wtab));
    }
    ctx->h_wtab = nullptr;
    ctx->h_wtab_cap = 0;
    CUDA_TRY(ctx, cudaMallocHost((void**)&ctx->h_wtab, n_w * 2 * sizeof(double)));
    ctx->h_wtab_cap = n_w * 2;
    ctx->wtab_flip = 0;
  }
  double* hw = ctx->h_wtab + (ctx->wtab_flip ? n_w : 0);
  if (ctx->wtab_in_flight[ctx->wtab_flip]) {
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_wtab[ctx->wtab_flip]));
    ctx->wtab_in_flight[ctx->wtab_flip] = false;
  }
  for (int q = 0; q < n_k; ++q) fill_wtab(ctx, B, q, xs[q], hw + (size_t)q * B.w_stride);
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_wtab, hw, n_w * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_wtab[ctx->wtab_flip], ctx->stream));
  ctx->wtab_in_flight[ctx->wtab_flip] = true;
  ctx->wtab_flip ^= 1;
  if (B.out_total > 0) {
    assemble_kernel<<<(unsigned)((B.out_total + ASM_NT - 1) / ASM_NT), ASM_NT, 0, ctx->stream>>>(
        B.d_asm_ptr, (const AsmEntry*)B.d_asm, ctx->d_wtab, raw_device, packed_device, B.out_total);
    ctx->last_launches++;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return ASLSD_OK;
}

static int64_t batch_out_len(const aslsd_ctx* ctx, int k) {
  const int d = ctx->d;
  int64_t out = 0;
  for (int i = 0; i < d; ++i) {
    const int np = ctx->kernels[k * d + i].n_param();
    out += (1 + np + 1) + 3 * (1 + np);
  }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j)
      if (i != j) out += 1 + ctx->kernels[k * d + i].n_param() + ctx->kernels[k * d + j].n_param();
  return out;
}

// upper bound of the number of rows of one sub-problem, independent of the parameters
static int64_t max_rows(const aslsd_ctx* ctx) {
  const int d = ctx->d;
  const int64_t R = (int64_t)ctx->kernels[d * d - 1].bases.size();
  int64_t items = 0;
  for (int i = 0; i < d; ++i) items += 2 * R + 2 * 2 * R * R;
  items += (int64_t)d * (d - 1) * R * R;
  return items;
}

static int check_batch_args(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs) {
  if (!ctx) return ASLSD_ERR_ARG;
  if (!ctx->has_model) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: no model set");
  if (n_k <= 0 || !ks || !xs) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: bad arguments");
  for (int q = 0; q < n_k; ++q)
    if (ks[q] < 0 || ks[q] >= ctx->d || !xs[q]) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: bad k");
  return ASLSD_OK;
}

// ---- C ABI --------------------------------------------------------------------------------------
extern "C" int aslsd_exact_out_len(const aslsd_ctx* ctx, int k, int64_t* out_len, int64_t* n_terms) {
  if (!ctx || !ctx->has_model || k < 0 || k >= ctx->d) return ASLSD_ERR_ARG;
  if (out_len) *out_len = batch_out_len(ctx, k);
  if (n_terms) *n_terms = 4 * ctx->d + ctx->d * (ctx->d - 1);
  return ASLSD_OK;
}

extern "C" int aslsd_set_option(aslsd_ctx* ctx, int option, int value) {
  if (!ctx) return ASLSD_ERR_ARG;
  if (option == ASLSD_OPT_LAG_QUADRATURE) { ctx->use_quadrature = value != 0; ctx->model_epoch++; return ASLSD_OK; }
  if (option == ASLSD_OPT_WEIGHT_CACHE) {
    ctx->cache_weights = value != 0;
    ctx->w_streams.clear();
    ctx->model_epoch++;
    return ASLSD_OK;
  }
  if (option == ASLSD_OPT_ASYNC_EVENTS) { ctx->async_events = value != 0; return ASLSD_OK; }
  if (option == ASLSD_OPT_DEVICE_ASSEMBLY) { ctx->device_assembly = value != 0; return ASLSD_OK; }
  return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_option: unknown option");
}

extern "C" int aslsd_exact_sums_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs, int flags,
                                      const int64_t* m_lo, const int64_t* m_hi, double* const* outs,
                                      int64_t* const* n_pair_terms) {
  int rc = check_batch_args(ctx, n_k, ks, xs);
  if (rc) return rc;
  if (!outs) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: bad arguments");
  BatchPlan* Bp = nullptr;
  if ((rc = get_batch(ctx, n_k, ks, xs, flags, m_lo, m_hi, true, &Bp))) return rc;
  BatchPlan& B = *Bp;
  const size_t raw_bytes = std::max<size_t>(1, B.rows.size()) * ASLSD_MAX_COMP * sizeof(double);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_raw, &ctx->raw_cap, raw_bytes))) return rc;
  if ((rc = run_batch(ctx, B, ctx->d_raw))) return rc;
  // Assembly: on the device once a row plan serves its second call (the iterations of a fit, the
  // steps of a benchmark: the map is built once and only the packed sums are read back); a plan
  // used once - and the weighted plans of marked models, which are - is assembled on the host
  // from the raw rows, which costs one traversal of the map either way.
  const bool on_device = ctx->device_assembly && !ctx->roles && (B.uses > 0 || B.asm_built);
  B.uses++;
  if (on_device) {
    int64_t total = 0;
    for (int q = 0; q < n_k; ++q) total += B.plans[q].out_len;
    const size_t bytes = std::max<int64_t>(1, total) * sizeof(double);
    if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_packed, &ctx->packed_cap, bytes))) return rc;
    if ((rc = assemble_device(ctx, B, n_k, xs, ctx->d_raw, ctx->d_packed))) return rc;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    if ((rc = ensure_pinned_buffer(ctx, bytes))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_packed, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = events_verdict(ctx, true))) return rc;
    for (int q = 0; q < n_k; ++q)
      std::memcpy(outs[q], (const double*)ctx->h_pinned + B.out_off[q], (size_t)B.plans[q].out_len * sizeof(double));
  } else {
    if ((rc = ensure_pinned_buffer(ctx, raw_bytes))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_raw, raw_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = events_verdict(ctx, true))) return rc;
    assemble_all(ctx, B, n_k, xs, (const double*)ctx->h_pinned, outs);
  }
  if (n_pair_terms)
    for (int q = 0; q < n_k && !rc; ++q)
      if (n_pair_terms[q]) rc = count_pair_terms(ctx, B.plans[q], B.lo, B.hi, n_pair_terms[q]);
  return rc;
}

// Multi-GPU form with the assembly on the device: the packed sums of the batch (sub-problems
// concatenated in call order, aslsd_exact_packed_len_batch doubles) are left in packed_device
// on the context's stream.  Assembly is linear, so the caller all-reduces these few thousand
// doubles instead of the raw rows, and only they are read back.
extern "C" int aslsd_exact_packed_len_batch(aslsd_ctx* ctx, int n_k, const int* ks, int64_t* packed_len) {
  if (!ctx || !ctx->has_model || n_k <= 0 || !ks || !packed_len) return ASLSD_ERR_ARG;
  int64_t n = 0;
  for (int q = 0; q < n_k; ++q) {
    if (ks[q] < 0 || ks[q] >= ctx->d) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: bad k");
    n += batch_out_len(ctx, ks[q]);
  }
  *packed_len = n;
  return ASLSD_OK;
}

extern "C" int aslsd_exact_packed_device_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs,
                                               int flags, const int64_t* m_lo, const int64_t* m_hi,
                                               double* packed_device) {
  int rc = check_batch_args(ctx, n_k, ks, xs);
  if (rc) return rc;
  if (!packed_device) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_packed_device: bad arguments");
  if (ctx->roles) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_packed_device: not for weighted plans");
  BatchPlan* Bp = nullptr;
  if ((rc = get_batch(ctx, n_k, ks, xs, flags, m_lo, m_hi, true, &Bp))) return rc;
  BatchPlan& B = *Bp;
  const size_t raw_bytes = std::max<size_t>(1, B.rows.size()) * ASLSD_MAX_COMP * sizeof(double);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_raw, &ctx->raw_cap, raw_bytes))) return rc;
  if ((rc = run_batch(ctx, B, ctx->d_raw))) return rc;
  B.uses++;
  if ((rc = assemble_device(ctx, B, n_k, xs, ctx->d_raw, packed_device))) return rc;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  return ASLSD_OK;
}

extern "C" int aslsd_exact_sums(aslsd_ctx* ctx, int k, const double* x_k, int flags, const int64_t* m_lo,
                                const int64_t* m_hi, double* out, int64_t* n_pair_terms) {
  if (!ctx || !x_k || !out) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_sums: bad arguments");
  const double* xs[1] = {x_k};
  double* outs[1] = {out};
  int64_t* cnts[1] = {n_pair_terms};
  return aslsd_exact_sums_batch(ctx, 1, &k, xs, flags, m_lo, m_hi, outs, n_pair_terms ? cnts : nullptr);
}

// Marked models (MTLH): the same terms with per-event weights.  roles[r * d + i] names the weight
// vector (index into aslsd_set_weights, -1 = none) applied to
//   r = 0: the sources of phi_i            (I_ki(xi^i_n),            mtlh_exact_estimator.py:838)
//   r = 1: the targets of upsilon_i_*      (I_ki(xi^i_m),            :896-897, :950)
//   r = 2: the sources of upsilon_*_i      (I_ki(xi^i_n))
//   r = 3: the targets of psi_i            (K_i term: I_ki(xi^i_m),  :707-753)
//   r = 4: the targets of upsilonzero_i    (I_ki(xi^i_m)^2,          :755-805)
// term_mask (may be NULL): one byte per term in output order, 0 = leave the term at zero.  The
// derivatives with respect to the impact parameters are value sums with one weight vector
// replaced by the derivative of the impact (host side: aslsd_b200/mtlh.py).
extern "C" int aslsd_exact_sums_weighted(aslsd_ctx* ctx, int k, const double* x_k, int flags,
                                         const int64_t* m_lo, const int64_t* m_hi, const int32_t* roles,
                                         const uint8_t* term_mask, double* out, int64_t* n_pair_terms) {
  if (!ctx || !x_k || !out || !roles) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_sums_weighted: bad arguments");
  if (!ctx->has_model) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: no model set");
  const int d = ctx->d;
  for (int r = 0; r < 5; ++r)
    for (int i = 0; i < d; ++i) {
      const int v = roles[r * d + i];
      if (v < 0) continue;
      if (v >= (int)ctx->weight_off.size() || ctx->weight_dim[v] != i)
        return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_sums_weighted: weight vector of another event type");
    }
  // plans built with weights are never reused for other calls
  ctx->model_epoch++;
  ctx->roles = roles;
  ctx->term_mask = term_mask;
  const int rc = aslsd_exact_sums(ctx, k, x_k, flags, m_lo, m_hi, out, n_pair_terms);
  ctx->roles = nullptr;
  ctx->term_mask = nullptr;
  ctx->model_epoch++;
  return rc;
}

// Split form for multi-GPU: device-side raw rows (summed across ranks by the caller's
// all-reduce) and host-side assembly.  The row table depends only on the model, the
// flags and the parameters - not on the target block - so it is the same on every rank.
extern "C" int aslsd_exact_raw_len_batch(aslsd_ctx* ctx, int n_k, int64_t* raw_len) {
  if (!ctx || !ctx->has_model || n_k <= 0) return ASLSD_ERR_ARG;
  if (raw_len) *raw_len = (int64_t)n_k * max_rows(ctx) * ASLSD_MAX_COMP;
  return ASLSD_OK;
}

// Largest lag any term of the batch can reach at these parameters: the halo a time-block shard
// needs before its first target (events further back contribute to none of its sums).  Needs
// the model only, not the events.
extern "C" int aslsd_exact_max_lag(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs, int flags,
                                   double* max_lag) {
  int rc = check_batch_args(ctx, n_k, ks, xs);
  if (rc) return rc;
  if (!max_lag) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_max_lag: bad arguments");
  BatchPlan* Bp = nullptr;
  if ((rc = get_batch(ctx, n_k, ks, xs, flags, nullptr, nullptr, false, &Bp))) return rc;
  double w = 0.;
  for (auto& it : Bp->rows)
    if (is_pair_family(it.family)) w = std::max(w, it.s_hi);
  *max_lag = w;
  return ASLSD_OK;
}

// Number of doubles of the raw buffer actually written by the last raw_device call (rows x
// ASLSD_MAX_COMP): the prefix to all-reduce / copy to the host.  Same on every rank.
extern "C" int aslsd_exact_raw_used(aslsd_ctx* ctx, int64_t* raw_used) {
  if (!ctx || !raw_used) return ASLSD_ERR_ARG;
  *raw_used = ctx->last_work[6] * ASLSD_MAX_COMP;
  return ASLSD_OK;
}

extern "C" int aslsd_exact_raw_len(aslsd_ctx* ctx, int k, int64_t* raw_len) {
  if (!ctx || !ctx->has_model || k < 0 || k >= ctx->d) return ASLSD_ERR_ARG;
  return aslsd_exact_raw_len_batch(ctx, 1, raw_len);
}

extern "C" int aslsd_exact_raw_device_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs,
                                            int flags, const int64_t* m_lo, const int64_t* m_hi,
                                            double* raw_device) {
  int rc = check_batch_args(ctx, n_k, ks, xs);
  if (rc) return rc;
  if (!raw_device) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_raw_device: bad arguments");
  BatchPlan* Bp = nullptr;
  if ((rc = get_batch(ctx, n_k, ks, xs, flags, m_lo, m_hi, true, &Bp))) return rc;
  return run_batch(ctx, *Bp, raw_device);
}

extern "C" int aslsd_exact_raw_device(aslsd_ctx* ctx, int k, const double* x_k, int flags, const int64_t* m_lo,
                                      const int64_t* m_hi, double* raw_device, int64_t* n_pair_terms) {
  if (!ctx || !x_k || !raw_device) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_raw_device: bad arguments");
  const double* xs[1] = {x_k};
  int rc = aslsd_exact_raw_device_batch(ctx, 1, &k, xs, flags, m_lo, m_hi, raw_device);
  if (rc || !n_pair_terms) return rc;
  BatchPlan* Bp = nullptr;
  if ((rc = get_batch(ctx, 1, &k, xs, flags, m_lo, m_hi, true, &Bp))) return rc;
  return count_pair_terms(ctx, Bp->plans[0], Bp->lo, Bp->hi, n_pair_terms);
}

extern "C" int aslsd_exact_assemble_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs,
                                          int flags, const double* raw_host, double* const* outs) {
  int rc = check_batch_args(ctx, n_k, ks, xs);
  if (rc) return rc;
  if (!raw_host || !outs) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_assemble: bad arguments");
  if ((rc = events_verdict(ctx, true))) return rc;   // raw_host is there: the device has long finished
  BatchPlan* Bp = nullptr;
  if ((rc = get_batch(ctx, n_k, ks, xs, flags, nullptr, nullptr, false, &Bp))) return rc;
  BatchPlan& B = *Bp;
  assemble_all(ctx, B, n_k, xs, raw_host, outs);
  return ASLSD_OK;
}

// The map of the device assembly (CSR table + omega tables) evaluated on the host with the
// kernel's arithmetic: a check of the table against aslsd_exact_assemble_batch that needs no device.
extern "C" int aslsd_exact_assemble_map_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs,
                                              int flags, const double* raw_host, double* const* outs) {
  int rc = check_batch_args(ctx, n_k, ks, xs);
  if (rc) return rc;
  if (!raw_host || !outs) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_assemble_map: bad arguments");
  BatchPlan* Bp = nullptr;
  if ((rc = get_batch(ctx, n_k, ks, xs, flags, nullptr, nullptr, false, &Bp))) return rc;
  BatchPlan& B = *Bp;
  std::vector<int64_t> ptr;
  std::vector<AsmEntry> entries;
  if ((rc = build_assembly_host(ctx, B, ptr, entries))) return rc;
  std::vector<double> w((size_t)n_k * B.w_stride);
  for (int q = 0; q < n_k; ++q) fill_wtab(ctx, B, q, xs[q], w.data() + (size_t)q * B.w_stride);
  for (int q = 0; q < n_k; ++q)
    for (int64_t c = 0; c < B.plans[q].out_len; ++c) {
      const int64_t o = B.out_off[q] + c;
      double acc = 0.;
      for (int64_t e = ptr[o]; e < ptr[o + 1]; ++e) {
        const AsmEntry& E = entries[(size_t)e];
        const double coef = ((w[E.a] * w[E.b]) * E.g) / E.gd;
        acc += w[E.o] * (coef * raw_host[E.raw_ix]);
      }
      outs[q][c] = acc;
    }
  return ASLSD_OK;
}

extern "C" int aslsd_exact_assemble(aslsd_ctx* ctx, int k, const double* x_k, int flags, const double* raw_host,
                                    double* out) {
  if (!ctx || !x_k || !raw_host || !out) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_assemble: bad arguments");
  const double* xs[1] = {x_k};
  double* outs[1] = {out};
  return aslsd_exact_assemble_batch(ctx, 1, &k, xs, flags, raw_host, outs);
}

// Work performed by the last exact-sum call (roofline accounting):
//   [0] basis-pair closed-form evaluations of the per-item pair kernels
//   [1] evaluations of the single-index kernels
//   [2] event pairs binned by the lag-moment kernel
//   [3] closed-form evaluations of the quadrature rows (row x node)
//   [4] quadrature streams, [5] bins over all streams
//   [6] rows (unique work items), [7] work items before de-duplication
extern "C" int aslsd_last_work_counts(aslsd_ctx* ctx, int64_t* counts, int n) {
  if (!ctx || !counts || n < 8) return ASLSD_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  // d_counts: [per-item work units | items | boundary groups]; d_qcounts: the moment kernels' work units
  const size_t n_old = (size_t)ctx->last_n_wu + (size_t)ctx->last_n_items;
  const size_t n_all = n_old + (size_t)ctx->last_n_bgroups;
  const size_t n_q = ctx->last_weights_reused ? 0 : (size_t)ctx->last_n_wu_q;
  std::vector<unsigned long long> h(std::max<size_t>(1, n_all)), hq(std::max<size_t>(1, n_q));
  if (n_all)
    CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), ctx->d_counts, sizeof(unsigned long long) * n_all,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  if (n_q)
    CUDA_TRY(ctx, cudaMemcpyAsync(hq.data(), ctx->d_qcounts, sizeof(unsigned long long) * n_q,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  if (n_all || n_q) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  int64_t pe = 0, se = 0;
  for (int f = 0; f < ITEM_N_FAMILIES; ++f) {
    const bool pair = is_pair_family(f);
    for (int x = 0; x < ctx->last_fam_count[f]; ++x)
      (pair ? pe : se) += (int64_t)h[(size_t)ctx->last_n_wu + ctx->last_fam_first[f] + x];
  }
  for (int64_t g = 0; g < ctx->last_n_bgroups; ++g) pe += (int64_t)h[n_old + (size_t)g];
  counts[0] = pe;
  counts[1] = se;
  int64_t pairs = 0;
  for (size_t w = 0; w < n_q; ++w) pairs += (int64_t)hq[w];
  counts[2] = pairs;
  for (int c = 3; c < 8; ++c) counts[c] = ctx->last_work[c];
  return ASLSD_OK;
}

// The rows of the last batch plan: what each one sums (include/aslsd_b200.h).  Host data only.
extern "C" int aslsd_exact_plan_rows(aslsd_ctx* ctx, int64_t cap, int64_t* n_rows, int32_t* meta, double* params) {
  if (!ctx || !n_rows || cap < 0 || (cap > 0 && (!meta || !params))) return ASLSD_ERR_ARG;
  const PlanCache* cache = (const PlanCache*)ctx->plan_cache;
  if (!cache || cache->plans.empty())
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "plan rows: no exact-sum call has been planned yet");
  const BatchPlan* B = cache->plans.front();
  *n_rows = (int64_t)B->rows.size();
  for (int64_t r = 0; r < std::min<int64_t>(cap, *n_rows); ++r) {
    const ItemDev& it = B->rows[(size_t)r];
    meta[7 * r + 0] = it.family;
    meta[7 * r + 1] = it.tgt_dim;
    meta[7 * r + 2] = it.src_dim;
    meta[7 * r + 3] = it.same;
    meta[7 * r + 4] = it.mask;
    // weight vectors of the targets / sources (marked models): index into aslsd_set_weights, -1 = none
    for (int side = 0; side < 2; ++side) {
      const int64_t off = side == 0 ? it.wa_off : it.wb_off;
      int32_t vec = -1;
      for (size_t v = 0; off >= 0 && v < ctx->weight_off.size(); ++v)
        if (ctx->weight_off[v] == off) vec = (int32_t)v;
      meta[7 * r + 5 + side] = vec;
    }
    const double v[11] = {it.b1, it.d1, it.b2, it.d2, it.bo, it.dlo, it.s_lo, it.s_hi, it.t_int,
                          (double)it.m_lo, (double)it.m_hi};
    for (int q = 0; q < 11; ++q) params[11 * r + q] = v[q];
  }
  return ASLSD_OK;
}

extern "C" int aslsd_last_eval_counts(aslsd_ctx* ctx, int64_t* pair_evals, int64_t* single_evals) {
  int64_t c[8];
  int rc = aslsd_last_work_counts(ctx, c, 8);
  if (rc) return rc;
  if (pair_evals) *pair_evals = c[0];
  if (single_evals) *single_evals = c[1];
  return ASLSD_OK;
}
