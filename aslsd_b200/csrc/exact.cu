// Exact LSE sums: batching over sub-problems, de-duplication of identical work
// items, class terms for the shared-class kernel, launches and assembly.
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
// Rows that are products of single-variable factors of the lag (Gaussian upsilon
// value / own-basis d omega, Gaussian phi value) are served, for their interior
// targets, by the shared-class kernel (class_kernel.cuh); their last few targets
// (T_f - t_m within the kernels' reach) and all other rows by the per-item pair
// kernels (pair_kernels.cuh).
#include "class_kernel.cuh"
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

struct RowKey {
  ItemDev it;
  bool operator<(const RowKey& o) const { return std::memcmp(&it, &o.it, sizeof(ItemDev)) < 0; }
};

struct BatchPlan {
  std::vector<double> key_rows, key_work;   // cache keys (see get_batch)
  bool has_work = false;
  bool device_resident = false;   // item / class tables of this plan are on the device
  int flags = 0;
  std::vector<int> ks;
  std::vector<SumPlan> plans;
  std::vector<std::vector<int64_t>> item_row;   // per plan: item -> row
  std::vector<ItemDev> rows;                    // canonical item of every row
  int64_t n_items_total = 0;                    // work items before de-duplication
  // per-item kernels
  std::vector<ItemDev> items;                   // device items, grouped by family
  int fam_first[8], fam_count[8];
  int64_t fam_wu_first[8], fam_wu_count[8];
  int64_t n_wu = 0;
  // shared-class kernel
  std::vector<ClassTermDev> cterms;
  std::vector<EClassDev> ecls;
  std::vector<NClassDev> ncls;
  std::vector<ProdDev> prods;
  int64_t n_wu_class = 0, n_class_partials = 0;
  size_t class_smem = 0;
};

static ItemDev canonical(const ItemDev& it) {
  ItemDev c = it;
  c.n_chunks = 0;
  c.tpw = 0;
  c.wu_first = 0;
  c.dst_row = 0;
  return c;
}

static bool class_eligible(const ItemDev& it) {
  if (it.family == ITEM_PHI_GAUSS) return it.mask == MASK(0);
  if (it.family != ITEM_UPS_GAUSS) return false;
  const int allowed = MASK(C_G) | MASK(C_SW) | MASK_OWN_SAME;
  if (it.mask & ~allowed) return false;
  if ((it.mask & MASK(C_SW)) && !(it.mask & MASK_OWN_SAME)) return false;
  return true;
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

template <class T, class Eq>
static int find_or_add(std::vector<T>& v, size_t first, const T& x, Eq eq) {
  for (size_t q = first; q < v.size(); ++q)
    if (eq(v[q], x)) return (int)(q - first);
  v.push_back(x);
  return (int)(v.size() - 1 - first);
}

static const double SQRT_2PI = 2.50662827463100050242;

// Turn the eligible rows of one (stream, target block) group into class terms.
static void build_class_terms(const aslsd_ctx* ctx, BatchPlan& B, const std::vector<int64_t>& group,
                              std::vector<bool>& fast, std::vector<int64_t>& boundary_from) {
  const ItemDev first = B.rows[group[0]];
  const bool is_phi = first.family == ITEM_PHI_GAUSS;
  double t_int = 0.;
  if (!is_phi)
    for (int64_t r : group) t_int = std::max(t_int, B.rows[r].t_int);
  const int64_t m_b = is_phi ? first.m_hi : interior_end(ctx, first.tgt_dim, first.m_lo, first.m_hi, t_int);
  if (m_b - first.m_lo < CK_TM / 4) return;   // not worth a launch: the per-item kernels take it all
  auto e_eq = [](const EClassDev& a, const EClassDev& b) { return a.dd == b.dd && a.isig == b.isig && a.pn == b.pn; };
  auto n_eq = [](const NClassDev& a, const NClassDev& b) { return a.one == b.one && a.c0 == b.c0 && a.c1 == b.c1; };
  size_t g = 0;
  while (g < group.size()) {
    ClassTermDev T;
    std::memset(&T, 0, sizeof(T));
    T.tgt_dim = first.tgt_dim;
    T.src_dim = first.src_dim;
    T.same = first.same;
    T.m_lo = first.m_lo;
    T.m_hi = m_b;
    T.e_first = (int32_t)B.ecls.size();
    T.n_first = (int32_t)B.ncls.size();
    T.prod_first = (int32_t)B.prods.size();
    T.s_lo = 1e300;
    T.s_hi = 0.;
    for (; g < group.size(); ++g) {
      const ItemDev it = B.rows[group[g]];
      // classes of this row
      EClassDev E;
      NClassDev N;
      std::memset(&E, 0, sizeof(E));
      std::memset(&N, 0, sizeof(N));
      if (is_phi) {
        E.dd = it.d1;
        E.isig = 1. / it.b1;
        E.pn = 1. / (it.b1 * SQRT_2PI);
        N.one = 1;
      } else {
        const double sig2 = it.b1 * it.b1 + it.b2 * it.b2, sig = std::sqrt(sig2);
        E.dd = it.d2 - it.d1;
        E.isig = 1. / sig;
        E.pn = 1. / (sig * SQRT_2PI);
        const double inv_bs = sig / (it.b1 * it.b2);
        const double r1 = (it.b2 * it.b2) / sig2, r2 = (it.b1 * it.b1) / sig2;
        // lower limit lo = c0 + c1 s of the reference's gaussian_int; stored divided by
        // sqrt(2): Q = 1 - Phi(lo) = erfc(lo / sqrt(2)) / 2
        N.c0 = -(r1 * it.d1 + r2 * it.d2) * inv_bs / std::sqrt(2.);
        N.c1 = r2 * inv_bs / std::sqrt(2.);
      }
      const size_t e_before = B.ecls.size(), n_before = B.ncls.size();
      const int e = find_or_add(B.ecls, (size_t)T.e_first, E, e_eq);
      const int n = find_or_add(B.ncls, (size_t)T.n_first, N, n_eq);
      int n_one = -1;
      const bool want_sw = !is_phi && (it.mask & MASK(C_SW));
      if (want_sw) {
        NClassDev one;
        std::memset(&one, 0, sizeof(one));
        one.one = 1;
        n_one = find_or_add(B.ncls, (size_t)T.n_first, one, n_eq);
      }
      const int n_new = 1 + (want_sw ? 1 : 0);
      const bool over = (int)(B.ecls.size() - T.e_first) > CK_MAX_CLASSES ||
                        (int)(B.ncls.size() - T.n_first) > CK_MAX_CLASSES ||
                        (int)(B.prods.size() - T.prod_first) + n_new > 32 * CK_MAXR;
      if (over && B.prods.size() > (size_t)T.prod_first) {
        B.ecls.resize(e_before);
        B.ncls.resize(n_before);
        break;   // close this class term, the row opens the next one
      }
      ProdDev P;
      std::memset(&P, 0, sizeof(P));
      P.e = (int16_t)e;
      P.n = (int16_t)n;
      P.dst = (int32_t)(group[g] * ASLSD_MAX_COMP + (is_phi ? 0 : C_G));
      P.scale = 1.;
      B.prods.push_back(P);
      if (want_sw) {
        // own-basis d/d omega: 2 P0(s) I0 with the verbatim (unshifted) limits; in the
        // interior I0 = 1 - Phi(-sqrt(2) delta / beta) is a constant of the basis
        const double lo0 = -std::sqrt(2.) * it.dlo / it.bo;
        const double i0 = 1.0 - 0.5 * std::erfc(-lo0 / std::sqrt(2.));
        P.n = (int16_t)n_one;
        P.dst = (int32_t)(group[g] * ASLSD_MAX_COMP + C_SW);
        P.scale = 2. * i0;
        B.prods.push_back(P);
      }
      T.s_lo = std::min(T.s_lo, it.s_lo);
      T.s_hi = std::max(T.s_hi, it.s_hi);
      fast[group[g]] = true;
      boundary_from[group[g]] = m_b;
    }
    T.nE = (int32_t)(B.ecls.size() - T.e_first);
    T.nN = (int32_t)(B.ncls.size() - T.n_first);
    T.nProd = (int32_t)(B.prods.size() - T.prod_first);
    B.class_smem = std::max(B.class_smem, class_smem_bytes(T.nE, T.nN, T.nProd));
    B.cterms.push_back(T);
  }
}

static int build_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs, int flags,
                       const int64_t* m_lo, const int64_t* m_hi, bool with_work, BatchPlan& B) {
  B.flags = flags;
  B.ks.assign(ks, ks + n_k);
  B.plans.resize(n_k);
  B.item_row.resize(n_k);
  std::map<RowKey, int64_t> index;
  for (int q = 0; q < n_k; ++q) {
    int rc = build_plan(ctx, ks[q], xs[q], flags, m_lo, m_hi, false, B.plans[q]);
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
  if (!with_work) return ASLSD_OK;
  const size_t n_rows = B.rows.size();
  std::vector<bool> fast(n_rows, false);
  std::vector<int64_t> boundary_from(n_rows, 0);
  if (ctx->use_class_kernel) {
    // group the eligible rows by (kind, streams, target block), in row order
    std::map<std::array<int64_t, 6>, std::vector<int64_t>> groups;
    std::vector<std::array<int64_t, 6>> order;
    for (size_t r = 0; r < n_rows; ++r) {
      const ItemDev& it = B.rows[r];
      if (!class_eligible(it) || it.m_hi <= it.m_lo) continue;
      std::array<int64_t, 6> key = {it.family, it.tgt_dim, it.src_dim, it.same, it.m_lo, it.m_hi};
      auto f = groups.find(key);
      if (f == groups.end()) {
        order.push_back(key);
        f = groups.emplace(key, std::vector<int64_t>()).first;
      }
      f->second.push_back((int64_t)r);
    }
    for (auto& key : order) build_class_terms(ctx, B, groups[key], fast, boundary_from);
  }
  // ---- per-item kernels: non-class rows entirely, class rows beyond the interior -----------
  std::vector<ItemDev> fam_items[8];
  for (size_t r = 0; r < n_rows; ++r) {
    ItemDev it = B.rows[r];
    it.dst_row = (int64_t)r;
    if (fast[r]) {
      if (boundary_from[r] >= it.m_hi) continue;
      it.m_lo = boundary_from[r];
    }
    fam_items[it.family].push_back(it);
  }
  int base = 0;
  int64_t wu = 0;
  for (int f = 0; f < 8; ++f) {
    B.fam_first[f] = base;
    B.fam_count[f] = (int)fam_items[f].size();
    B.fam_wu_first[f] = wu;
    const bool pair = (f == ITEM_UPS_EXP || f == ITEM_UPS_GAUSS || f == ITEM_PHI_EXP || f == ITEM_PHI_GAUSS);
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
      it.wu_first = wu;
      wu += it.n_chunks;
      B.items.push_back(it);
    }
    B.fam_wu_count[f] = wu - B.fam_wu_first[f];
    base += B.fam_count[f];
  }
  B.n_wu = wu;
  // ---- class terms: work units of whole tiles, a few waves over the 148 SMs ----------------
  int64_t tot = 0;
  for (auto& T : B.cterms) tot += T.m_hi - T.m_lo;
  int64_t tpw = (tot + 148 * 2 * 6 - 1) / (148 * 2 * 6);
  tpw = std::max<int64_t>(CK_TM, std::min<int64_t>(tpw, 32 * CK_TM));
  tpw = ((tpw + CK_TM - 1) / CK_TM) * CK_TM;
  int64_t cwu = 0, part = 0;
  for (auto& T : B.cterms) {
    T.tpw = tpw;
    T.n_chunks = (int32_t)((T.m_hi - T.m_lo + tpw - 1) / tpw);
    T.wu_first = cwu;
    T.part_first = part;
    cwu += T.n_chunks;
    part += (int64_t)T.n_chunks * T.nProd;
  }
  B.n_wu_class = cwu;
  B.n_class_partials = part;
  if (B.n_wu >= (int64_t)2147483647 || cwu >= (int64_t)2147483647)
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: too many work units");
  return ASLSD_OK;
}

// Plans are pure functions of (model, events, flags, target block, non-omega
// parameters): the last one is kept and reused, with the omega-dependent variables
// refreshed.  No result is cached - every call launches all of its kernels.
static void free_batch(void* p) { delete (BatchPlan*)p; }

static int get_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs, int flags,
                     const int64_t* m_lo, const int64_t* m_hi, bool with_work, BatchPlan** out) {
  std::vector<double> key_rows, key_work;
  key_rows.push_back((double)ctx->epoch);
  key_rows.push_back((double)n_k);
  key_rows.push_back((double)flags);
  for (int q = 0; q < n_k; ++q) {
    key_rows.push_back((double)ks[q]);
    plan_shape_key(ctx, ks[q], xs[q], key_rows);
  }
  key_work.push_back(ctx->use_class_kernel ? 1. : 0.);
  for (int i = 0; i < ctx->d; ++i) {
    key_work.push_back(m_lo ? (double)m_lo[i] : -1.);
    key_work.push_back(m_hi ? (double)m_hi[i] : -1.);
  }
  BatchPlan* C = (BatchPlan*)ctx->plan_cache;
  if (C && C->key_rows == key_rows && (!with_work || (C->has_work && C->key_work == key_work))) {
    for (int q = 0; q < n_k; ++q) {
      int rc = plan_vars(ctx, ks[q], xs[q], C->plans[q]);
      if (rc) return rc;
    }
    *out = C;
    return ASLSD_OK;
  }
  BatchPlan* B = new BatchPlan();
  int rc = build_batch(ctx, n_k, ks, xs, flags, m_lo, m_hi, with_work, *B);
  if (rc) { delete B; return rc; }
  B->key_rows.swap(key_rows);
  B->key_work.swap(key_work);
  B->has_work = with_work;
  if (ctx->plan_cache) free_batch(ctx->plan_cache);
  ctx->plan_cache = B;
  ctx->plan_cache_free = free_batch;
  *out = B;
  return ASLSD_OK;
}

template <int FAMILY>
static void launch_pair(aslsd_ctx* ctx, const BatchPlan& B, const ItemDev* d_items) {
  if (B.fam_wu_count[FAMILY] == 0) return;
  pair_kernel<FAMILY><<<(unsigned)B.fam_wu_count[FAMILY], PAIR_NT, 0, ctx->stream>>>(
      (const DimDev*)ctx->d_dims, d_items + B.fam_first[FAMILY], B.fam_count[FAMILY], ctx->T_f,
      ctx->d_partials, ctx->d_counts, B.fam_wu_first[FAMILY]);
  ctx->last_launches++;
}
template <int FAMILY>
static void launch_single(aslsd_ctx* ctx, const BatchPlan& B, const ItemDev* d_items) {
  if (B.fam_wu_count[FAMILY] == 0) return;
  single_kernel<FAMILY><<<(unsigned)B.fam_wu_count[FAMILY], PAIR_NT, 0, ctx->stream>>>(
      (const DimDev*)ctx->d_dims, d_items + B.fam_first[FAMILY], B.fam_count[FAMILY], ctx->T_f,
      ctx->d_partials, ctx->d_counts, B.fam_wu_first[FAMILY]);
  ctx->last_launches++;
}

// Launch everything; leaves raw[row][comp] in raw_device (zeroed here first).
static int run_batch(aslsd_ctx* ctx, BatchPlan& B, double* raw_device) {
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t n_items = B.items.size(), n_rows = B.rows.size();
  int rc;
  ctx->last_launches = 0;
  ctx->last_n_wu = B.n_wu;
  ctx->last_n_items = (int)n_items;
  ctx->last_n_wu_class = B.n_wu_class;
  for (int f = 0; f < 8; ++f) { ctx->last_fam_first[f] = B.fam_first[f]; ctx->last_fam_count[f] = B.fam_count[f]; }
  ctx->last_work[6] = (int64_t)n_rows;
  ctx->last_work[7] = B.n_items_total;
  if ((rc = ensure_device_buffer(ctx, &ctx->d_items, &ctx->items_cap, std::max<size_t>(1, n_items) * sizeof(ItemDev))))
    return rc;
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_partials, &ctx->partials_cap,
                                 std::max<size_t>(1, (size_t)B.n_wu) * ASLSD_MAX_COMP * sizeof(double)))) return rc;
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_counts, &ctx->counts_cap,
                                 ((size_t)B.n_wu + n_items + (size_t)B.n_wu_class + 1) * sizeof(unsigned long long))))
    return rc;
  if (n_items && !B.device_resident)
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_items, B.items.data(), n_items * sizeof(ItemDev),
                                  cudaMemcpyHostToDevice, ctx->stream));
  // class-kernel tables: one block [terms | P classes | Q classes | products]
  const size_t o_t = 0;
  const size_t o_e = o_t + ((B.cterms.size() * sizeof(ClassTermDev) + 15) & ~(size_t)15);
  const size_t o_n = o_e + ((B.ecls.size() * sizeof(EClassDev) + 15) & ~(size_t)15);
  const size_t o_p = o_n + ((B.ncls.size() * sizeof(NClassDev) + 15) & ~(size_t)15);
  const size_t c_bytes = o_p + B.prods.size() * sizeof(ProdDev);
  if (!B.cterms.empty() && B.device_resident) {
    if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_class_partials, &ctx->class_partials_cap,
                                   (size_t)B.n_class_partials * sizeof(double)))) return rc;
  }
  if (!B.cterms.empty() && !B.device_resident) {
    if ((rc = ensure_device_buffer(ctx, &ctx->d_class, &ctx->class_cap, c_bytes))) return rc;
    if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_class_partials, &ctx->class_partials_cap,
                                   (size_t)B.n_class_partials * sizeof(double)))) return rc;
    if ((rc = ensure_pinned_buffer(ctx, c_bytes))) return rc;
    char* hp = (char*)ctx->h_pinned;
    std::memcpy(hp + o_t, B.cterms.data(), B.cterms.size() * sizeof(ClassTermDev));
    std::memcpy(hp + o_e, B.ecls.data(), B.ecls.size() * sizeof(EClassDev));
    std::memcpy(hp + o_n, B.ncls.data(), B.ncls.size() * sizeof(NClassDev));
    std::memcpy(hp + o_p, B.prods.data(), B.prods.size() * sizeof(ProdDev));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_class, hp, c_bytes, cudaMemcpyHostToDevice, ctx->stream));
    // the pinned staging block is reused by the caller once the copy has completed
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  B.device_resident = true;
  // static per-pair costs of the class terms are turned into totals in aslsd_last_work_counts
  ctx->last_class_cost.clear();
  for (auto& T : B.cterms) {
    int n_norm = 0;
    for (int q = 0; q < T.nN; ++q) n_norm += B.ncls[T.n_first + q].one ? 0 : 1;
    for (int w = 0; w < T.n_chunks; ++w) ctx->last_class_cost.push_back({T.nE, n_norm, T.nProd});
  }
  const ItemDev* d_items = (const ItemDev*)ctx->d_items;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(raw_device, 0, std::max<size_t>(1, n_rows) * ASLSD_MAX_COMP * sizeof(double),
                                ctx->stream));
  if (!B.cterms.empty()) {
    const char* dc = (const char*)ctx->d_class;
    if (B.class_smem > ctx->class_smem_set) {
      CUDA_TRY(ctx, cudaFuncSetAttribute(class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)B.class_smem));
      ctx->class_smem_set = B.class_smem;
    }
    class_kernel<<<(unsigned)B.n_wu_class, CK_NT, B.class_smem, ctx->stream>>>(
        (const DimDev*)ctx->d_dims, (const ClassTermDev*)(dc + o_t), (int)B.cterms.size(),
        (const EClassDev*)(dc + o_e), (const NClassDev*)(dc + o_n), (const ProdDev*)(dc + o_p),
        ctx->d_class_partials, ctx->d_counts + B.n_wu + n_items);
    ctx->last_launches++;
  }
  launch_pair<ITEM_UPS_GAUSS>(ctx, B, d_items);
  launch_pair<ITEM_UPS_EXP>(ctx, B, d_items);
  launch_pair<ITEM_PHI_GAUSS>(ctx, B, d_items);
  launch_pair<ITEM_PHI_EXP>(ctx, B, d_items);
  launch_single<ITEM_UPS0_GAUSS>(ctx, B, d_items);
  launch_single<ITEM_UPS0_EXP>(ctx, B, d_items);
  launch_single<ITEM_PSI_GAUSS>(ctx, B, d_items);
  launch_single<ITEM_PSI_EXP>(ctx, B, d_items);
  if (n_items) {
    const int n_fin = (int)n_items * (ASLSD_MAX_COMP + 1);
    finalize_kernel<<<(n_fin + 127) / 128, 128, 0, ctx->stream>>>(d_items, (int)n_items, ctx->d_partials,
                                                                  ctx->d_counts, raw_device,
                                                                  ctx->d_counts + B.n_wu);
    ctx->last_launches++;
  }
  if (!B.cterms.empty()) {
    const char* dc = (const char*)ctx->d_class;
    class_finalize_kernel<<<(unsigned)B.cterms.size(), 128, 0, ctx->stream>>>(
        (const ClassTermDev*)(dc + o_t), (int)B.cterms.size(), (const ProdDev*)(dc + o_p),
        ctx->d_class_partials, raw_device);
    ctx->last_launches++;
  }
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timing_pending = true;
  CUDA_TRY(ctx, cudaGetLastError());
  return ASLSD_OK;
}

static int count_pair_terms(aslsd_ctx* ctx, const SumPlan& P, int64_t* n_pair_terms) {
  const size_t nt = P.terms.size();
  unsigned long long* d_cnt = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&d_cnt, nt * sizeof(unsigned long long)));
  cudaMemsetAsync(d_cnt, 0, nt * sizeof(unsigned long long), ctx->stream);
  for (size_t t = 0; t < nt; ++t) {
    const TermPlan& T = P.terms[t];
    if (T.kind < 2 || P.term_s_hi[t] < 0.) continue;
    int tgt = T.kind == 2 ? P.k : T.i;
    int src = T.kind == 2 ? T.i : T.j;
    // target block of the term: recover from any of its items
    int64_t m_lo = 0, m_hi = 0;
    int n_slots = (T.kind == 2) ? (int)ctx->kernels[P.k * ctx->d + T.i].bases.size()
                                : (int)(ctx->kernels[0].bases.size() * ctx->kernels[0].bases.size());
    for (int s = 0; s < n_slots; ++s) {
      int idx = P.slot_item[P.term_val_off[t] + s];
      if (idx >= 0) { m_lo = P.items[idx].m_lo; m_hi = P.items[idx].m_hi; break; }
    }
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
  if (option == ASLSD_OPT_CLASS_KERNEL) { ctx->use_class_kernel = value != 0; ctx->epoch++; return ASLSD_OK; }
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
  if ((rc = ensure_pinned_buffer(ctx, raw_bytes))) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_raw, raw_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  for (int q = 0; q < n_k; ++q)
    assemble(ctx, B.plans[q], B.item_row[q], xs[q], (const double*)ctx->h_pinned, outs[q]);
  if (n_pair_terms)
    for (int q = 0; q < n_k && !rc; ++q)
      if (n_pair_terms[q]) rc = count_pair_terms(ctx, B.plans[q], n_pair_terms[q]);
  return rc;
}

extern "C" int aslsd_exact_sums(aslsd_ctx* ctx, int k, const double* x_k, int flags, const int64_t* m_lo,
                                const int64_t* m_hi, double* out, int64_t* n_pair_terms) {
  if (!ctx || !x_k || !out) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_sums: bad arguments");
  const double* xs[1] = {x_k};
  double* outs[1] = {out};
  int64_t* cnts[1] = {n_pair_terms};
  return aslsd_exact_sums_batch(ctx, 1, &k, xs, flags, m_lo, m_hi, outs, n_pair_terms ? cnts : nullptr);
}

// Split form for multi-GPU: device-side raw rows (summed across ranks by the caller's
// all-reduce) and host-side assembly.  The row table depends only on the model, the
// flags and the parameters - not on the target block - so it is the same on every rank.
extern "C" int aslsd_exact_raw_len_batch(aslsd_ctx* ctx, int n_k, int64_t* raw_len) {
  if (!ctx || !ctx->has_model || n_k <= 0) return ASLSD_ERR_ARG;
  if (raw_len) *raw_len = (int64_t)n_k * max_rows(ctx) * ASLSD_MAX_COMP;
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
  SumPlan P;
  if ((rc = build_plan(ctx, k, x_k, flags, m_lo, m_hi, false, P))) return rc;
  return count_pair_terms(ctx, P, n_pair_terms);
}

extern "C" int aslsd_exact_assemble_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs,
                                          int flags, const double* raw_host, double* const* outs) {
  int rc = check_batch_args(ctx, n_k, ks, xs);
  if (rc) return rc;
  if (!raw_host || !outs) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_assemble: bad arguments");
  BatchPlan* Bp = nullptr;
  if ((rc = get_batch(ctx, n_k, ks, xs, flags, nullptr, nullptr, false, &Bp))) return rc;
  BatchPlan& B = *Bp;
  for (int q = 0; q < n_k; ++q) assemble(ctx, B.plans[q], B.item_row[q], xs[q], raw_host, outs[q]);
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
//   [2] event pairs visited by the shared-class kernel
//   [3] exp evaluations, [4] normcdf evaluations, [5] multiply-adds of the shared-class kernel
//   [6] rows (unique work items), [7] work items before de-duplication
extern "C" int aslsd_last_work_counts(aslsd_ctx* ctx, int64_t* counts, int n) {
  if (!ctx || !counts || n < 8) return ASLSD_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t n_old = (size_t)ctx->last_n_wu + (size_t)ctx->last_n_items;
  const size_t n_all = n_old + (size_t)ctx->last_n_wu_class;
  std::vector<unsigned long long> h(std::max<size_t>(1, n_all));
  if (n_all) {
    CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), ctx->d_counts, sizeof(unsigned long long) * n_all,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  int64_t pe = 0, se = 0;
  for (int f = 0; f < 8; ++f) {
    const bool pair = (f == ITEM_UPS_EXP || f == ITEM_UPS_GAUSS || f == ITEM_PHI_EXP || f == ITEM_PHI_GAUSS);
    for (int x = 0; x < ctx->last_fam_count[f]; ++x)
      (pair ? pe : se) += (int64_t)h[(size_t)ctx->last_n_wu + ctx->last_fam_first[f] + x];
  }
  counts[0] = pe;
  counts[1] = se;
  int64_t pairs = 0, n_exp = 0, n_norm = 0, n_fma = 0;
  for (int64_t w = 0; w < ctx->last_n_wu_class; ++w) {
    const int64_t p = (int64_t)h[n_old + (size_t)w];
    pairs += p;
    if ((size_t)w < ctx->last_class_cost.size()) {
      n_exp += p * ctx->last_class_cost[w][0];
      n_norm += p * ctx->last_class_cost[w][1];
      n_fma += p * ctx->last_class_cost[w][2];
    }
  }
  counts[2] = pairs;
  counts[3] = n_exp;
  counts[4] = n_norm;
  counts[5] = n_fma;
  counts[6] = ctx->last_work[6];
  counts[7] = ctx->last_work[7];
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
