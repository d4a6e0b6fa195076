// Exact LSE sums of one sub-problem k: planning (terms -> work items), launches
// and assembly.  Replaces MHPExactEstim.compute_objective and its five
// compute_*_sum methods (optimize/estimators/mhp_exact_estimator.py:543-761).
#include "pair_kernels.cuh"

#include <array>
#include <cmath>
#include <cstring>

typedef std::array<double, 3> Vars;

// Effective support of the kernels: lags outside contribute < 1e-17 of a term
// (agreement with the untruncated sum ~1e-15).
static const double Z_GAUSS = 9.0;     // Gaussian half-width in standard deviations
static const double C_EXP = 42.0;      // exponential support in units of 1/beta
static const double X_INTERIOR = 10.0; // upper integration limit treated as +inf beyond this
static const double C_INTERIOR = 60.0; // exp(-x) treated as 0 beyond this

struct ExactPlan {
  int k = -1, flags = 0;
  std::vector<TermPlan> terms;
  std::vector<ItemDev> items;       // grouped by family
  int fam_first[8], fam_count[8];
  int64_t fam_wu_first[8], fam_wu_count[8];
  int64_t n_wu = 0;
  // per-term tables of item indices (-1: empty window, contributes zero)
  std::vector<int> slot_item;
  std::vector<int> term_val_off, term_der_off;  // offsets into slot_item
  int64_t out_len = 0;
  std::vector<std::vector<Vars>> V;   // per source dim i: vars of K[k][i] with x_ki
  std::vector<std::vector<Vars>> VU;  // per source dim i: vars of KU with x_ki
  std::vector<double> term_s_lo, term_s_hi;  // union window per term (pair counting)
};

static void kernel_vars(const KernelHost& K, const double* params, std::vector<Vars>& V) {
  V.resize(K.bases.size());
  int pos = 0;
  for (size_t r = 0; r < K.bases.size(); ++r) {
    const BasisHost& b = K.bases[r];
    V[r] = Vars{0., 0., 0.};
    for (int v = 0; v < b.n_vars(); ++v) V[r][v] = ((b.free_mask >> v) & 1) ? params[pos++] : b.fixed[v];
  }
}

static bool vars_equal(const std::vector<Vars>& a, const std::vector<Vars>& b) {
  if (a.size() != b.size()) return false;
  for (size_t r = 0; r < a.size(); ++r)
    for (int v = 1; v < 3; ++v)   // omega does not enter the raw components
      if (a[r][v] != b[r][v]) return false;
  return true;
}

static void ups_window(int type, const Vars& v1, const Vars& v2, double* s_lo, double* s_hi, double* t_int) {
  if (type == ASLSD_BASIS_EXPONENTIAL) {
    *s_lo = 0.;
    *s_hi = C_EXP / v2[1];
    *t_int = C_INTERIOR / (v1[1] + v2[1]);
  } else {
    const double sig = std::sqrt(v1[1] * v1[1] + v2[1] * v2[1]);
    const double dd = v2[2] - v1[2];
    *s_lo = std::max(0., dd - Z_GAUSS * sig);
    *s_hi = dd + Z_GAUSS * sig;
    const double bs = v1[1] * v2[1] / sig;
    const double r1 = v2[1] * v2[1] / (sig * sig), r2 = v1[1] * v1[1] / (sig * sig);
    *t_int = X_INTERIOR * bs + r1 * v1[2] + r2 * std::max(v2[2], 0.);
  }
}

static void own_window(int type, const Vars& v, double* s_hi, double* t_int) {
  if (type == ASLSD_BASIS_EXPONENTIAL) {
    *s_hi = C_EXP / v[1];
    *t_int = C_INTERIOR / (2. * v[1]);
  } else {
    *s_hi = Z_GAUSS * std::sqrt(2.) * v[1];
    *t_int = v[2] + X_INTERIOR * v[1] / std::sqrt(2.);
  }
}

struct ItemRef { int fam, idx; };

// Build the work items of an upsilon-type term.
//   VA1/VA2: variables of the value closure (structure KU), VB: variables of the
//   derivative closure of a same-type term (K[k][i]); merged when equal.
static void plan_upsilon(const aslsd_ctx* ctx, ExactPlan& P, std::vector<ItemDev> fam_items[8],
                         std::vector<std::pair<int, ItemRef>>& fixups, const KernelHost& KU,
                         const std::vector<Vars>& VA1, const std::vector<Vars>& VA2,
                         const KernelHost* KB, const std::vector<Vars>* VB, int tgt, int src, bool single,
                         int64_t m_lo, int64_t m_hi, bool diff, double* u_lo, double* u_hi) {
  const int R1 = (int)KU.bases.size(), R2 = R1;
  const bool sim = (KB != nullptr);
  const bool merged = sim && vars_equal(VA1, *VB);
  const int type = KU.bases[0].type;
  const int fam = single ? (type == ASLSD_BASIS_EXPONENTIAL ? ITEM_UPS0_EXP : ITEM_UPS0_GAUSS)
                         : (type == ASLSD_BASIS_EXPONENTIAL ? ITEM_UPS_EXP : ITEM_UPS_GAUSS);
  *u_lo = 1e300;
  *u_hi = 0.;
  auto emit = [&](int p, int q, const Vars& v1, const Vars& v2, int mask, bool own, const Vars& vo) {
    ItemDev it;
    std::memset(&it, 0, sizeof(it));
    it.family = fam;
    it.tgt_dim = tgt;
    it.src_dim = src;
    it.same = (tgt == src) ? 1 : 0;
    it.m_lo = m_lo;
    it.m_hi = m_hi;
    it.b1 = v1[1]; it.d1 = v1[2]; it.b2 = v2[1]; it.d2 = v2[2];
    it.bo = vo[1]; it.dlo = vo[2];
    ups_window(type, v1, v2, &it.s_lo, &it.s_hi, &it.t_int);
    if (own) {
      double oh, ot;
      own_window(type, vo, &oh, &ot);
      it.s_lo = 0.;
      it.s_hi = std::max(it.s_hi, oh);
      it.t_int = std::max(it.t_int, ot);
      if (vo[1] == v1[1] && vo[1] == v2[1] && vo[2] == v1[2] && vo[2] == v2[2]) mask |= MASK_OWN_SAME;
    }
    it.mask = mask;
    if (it.s_hi < 0.) return -1;   // empty lag window: contributes exactly zero
    *u_lo = std::min(*u_lo, it.s_lo);
    *u_hi = std::max(*u_hi, it.s_hi);
    fam_items[fam].push_back(it);
    return (int)fam_items[fam].size() - 1;
  };
  const std::vector<BasisHost>& BU = KU.bases;
  // value set (and, when merged or for cross terms, the derivative components too)
  P.term_val_off.push_back((int)P.slot_item.size());
  for (int p = 0; p < R1; ++p)
    for (int q = 0; q < R2; ++q) {
      int mask = MASK(C_G);
      bool own = false;
      if (diff && !sim) {
        if ((BU[p].free_mask >> 1) & 1) mask |= MASK(C_B1);
        if ((BU[p].free_mask >> 2) & 1) mask |= MASK(C_D1);
        if ((BU[q].free_mask >> 1) & 1) mask |= MASK(C_B2);
        if ((BU[q].free_mask >> 2) & 1) mask |= MASK(C_D2);
      }
      if (diff && merged) {
        const std::vector<BasisHost>& BB = KB->bases;
        if (p != q) {
          if ((BB[p].free_mask >> 1) & 1) mask |= MASK(C_B1);
          if ((BB[p].free_mask >> 2) & 1) mask |= MASK(C_D1);
          if ((BB[q].free_mask >> 1) & 1) mask |= MASK(C_B2);
          if ((BB[q].free_mask >> 2) & 1) mask |= MASK(C_D2);
        } else {
          if ((BB[p].free_mask >> 0) & 1) mask |= MASK(C_SW);
          if ((BB[p].free_mask >> 1) & 1) mask |= MASK(C_SB);
          if ((BB[p].free_mask >> 2) & 1) mask |= MASK(C_SD);
          own = (mask & (MASK(C_SW) | MASK(C_SB) | MASK(C_SD))) != 0;
        }
      }
      int idx = emit(p, q, VA1[p], VA2[q], mask, own, own ? (*VB)[p] : VA1[p]);
      fixups.push_back({(int)P.slot_item.size(), ItemRef{fam, idx}});
      P.slot_item.push_back(idx);
    }
  // separate derivative set of a same-type term whose own kernel differs from KU
  P.term_der_off.push_back((int)P.slot_item.size());
  if (sim && !merged && diff) {
    const std::vector<BasisHost>& BB = KB->bases;
    for (int p = 0; p < R1; ++p)
      for (int q = 0; q < R2; ++q) {
        int mask = MASK(C_G);
        bool own = false;
        if (p != q) {
          if ((BB[p].free_mask >> 1) & 1) mask |= MASK(C_B1);
          if ((BB[p].free_mask >> 2) & 1) mask |= MASK(C_D1);
          if ((BB[q].free_mask >> 1) & 1) mask |= MASK(C_B2);
          if ((BB[q].free_mask >> 2) & 1) mask |= MASK(C_D2);
        } else {
          if ((BB[p].free_mask >> 0) & 1) mask |= MASK(C_SW);
          if ((BB[p].free_mask >> 1) & 1) mask |= MASK(C_SB);
          if ((BB[p].free_mask >> 2) & 1) mask |= MASK(C_SD);
          own = (mask & (MASK(C_SW) | MASK(C_SB) | MASK(C_SD))) != 0;
        }
        int idx = emit(p, q, (*VB)[p], (*VB)[q], mask, own, (*VB)[p]);
        fixups.push_back({(int)P.slot_item.size(), ItemRef{fam, idx}});
        P.slot_item.push_back(idx);
      }
  } else if (sim && merged) {
    P.term_der_off.back() = P.term_val_off.back();
  }
  if (*u_lo > *u_hi) { *u_lo = 0.; *u_hi = -1.; }
}

static void plan_basis_items(ExactPlan& P, std::vector<ItemDev> fam_items[8],
                             std::vector<std::pair<int, ItemRef>>& fixups, const KernelHost& K,
                             const std::vector<Vars>& V, bool is_phi, int tgt, int src, int64_t m_lo,
                             int64_t m_hi, bool diff, double* u_lo, double* u_hi) {
  P.term_val_off.push_back((int)P.slot_item.size());
  *u_lo = 1e300;
  *u_hi = 0.;
  for (size_t r = 0; r < K.bases.size(); ++r) {
    const BasisHost& b = K.bases[r];
    const bool is_exp = b.type == ASLSD_BASIS_EXPONENTIAL;
    ItemDev it;
    std::memset(&it, 0, sizeof(it));
    it.family = is_phi ? (is_exp ? ITEM_PHI_EXP : ITEM_PHI_GAUSS) : (is_exp ? ITEM_PSI_EXP : ITEM_PSI_GAUSS);
    it.tgt_dim = tgt;
    it.src_dim = src;
    it.same = (tgt == src) ? 1 : 0;
    it.m_lo = m_lo;
    it.m_hi = m_hi;
    it.b1 = V[r][1];
    it.d1 = V[r][2];
    it.b2 = it.b1; it.d2 = it.d1; it.bo = it.b1; it.dlo = it.d1;
    int mask = MASK(0);
    if (diff) {
      if ((b.free_mask >> 1) & 1) mask |= MASK(1);
      if (!is_exp && ((b.free_mask >> 2) & 1)) mask |= MASK(2);
    }
    it.mask = mask;
    if (is_phi) {
      if (is_exp) { it.s_lo = 0.; it.s_hi = C_EXP / V[r][1]; }
      else { it.s_lo = std::max(0., V[r][2] - Z_GAUSS * V[r][1]); it.s_hi = V[r][2] + Z_GAUSS * V[r][1]; }
      *u_lo = std::min(*u_lo, it.s_lo);
      *u_hi = std::max(*u_hi, it.s_hi);
    }
    it.t_int = 1e300;   // psi is evaluated per target
    fam_items[it.family].push_back(it);
    const int idx = (int)fam_items[it.family].size() - 1;
    fixups.push_back({(int)P.slot_item.size(), ItemRef{it.family, idx}});
    P.slot_item.push_back(idx);
  }
  P.term_der_off.push_back(P.term_val_off.back());
}

static int build_plan(aslsd_ctx* ctx, int k, const double* x_k, int flags, const int64_t* m_lo,
                      const int64_t* m_hi, ExactPlan& P) {
  const int d = ctx->d;
  if (!ctx->has_model) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: no model set");
  if (!ctx->d_times) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: no events set");
  if (k < 0 || k >= d) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: bad k");
  P.k = k;
  P.flags = flags;
  const bool diff = (flags & ASLSD_DIFF) != 0;
  const KernelHost& KU = ctx->kernels[d * d - 1];
  std::vector<int64_t> lo(d), hi(d);
  for (int i = 0; i < d; ++i) {
    lo[i] = m_lo ? std::max<int64_t>(0, m_lo[i]) : 0;
    hi[i] = m_hi ? std::min<int64_t>(ctx->n_events[i], m_hi[i]) : ctx->n_events[i];
  }
  // parameter slices (mhp.py:276-299 interval_map)
  P.V.resize(d);
  P.VU.resize(d);
  std::vector<int> par_off(d + 1, 1);
  for (int i = 0; i < d; ++i) par_off[i + 1] = par_off[i] + ctx->kernels[k * d + i].n_param();
  for (int i = 0; i < d; ++i) {
    kernel_vars(ctx->kernels[k * d + i], x_k + par_off[i], P.V[i]);
    kernel_vars(KU, x_k + par_off[i], P.VU[i]);
    for (auto& v : P.V[i])
      if (!(v[1] > 0.)) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: beta must be positive");
  }
  std::vector<ItemDev> fam_items[8];
  std::vector<std::pair<int, ItemRef>> fixups;
  int64_t out = 0;
  auto add_term = [&](int kind, int i, int j, int n1, int n2) {
    TermPlan t;
    t.kind = kind; t.i = i; t.j = j; t.n1 = n1; t.n2 = n2; t.out_off = out;
    t.item_first = 0; t.item_count = 0;
    out += 1 + n1 + n2;
    P.terms.push_back(t);
  };
  double ul, uh;
  for (int i = 0; i < d; ++i) {
    const KernelHost& Kki = ctx->kernels[k * d + i];
    const int np = Kki.n_param();
    // psi_i (:543-578)
    add_term(0, i, i, np, 1);
    plan_basis_items(P, fam_items, fixups, Kki, P.V[i], false, i, i, lo[i], hi[i], diff, &ul, &uh);
    P.term_s_lo.push_back(0.); P.term_s_hi.push_back(-1.);
    // upsilonzero_i (:580-609)
    add_term(1, i, i, np, 0);
    plan_upsilon(ctx, P, fam_items, fixups, KU, P.VU[i], P.VU[i], &Kki, &P.V[i], i, i, true, lo[i], hi[i], diff,
                 &ul, &uh);
    P.term_s_lo.push_back(0.); P.term_s_hi.push_back(-1.);
    // phi_i (:611-638): targets of type k, sources of type i
    add_term(2, i, i, np, 0);
    plan_basis_items(P, fam_items, fixups, Kki, P.V[i], true, k, i, lo[k], hi[k], diff, &ul, &uh);
    P.term_s_lo.push_back(ul); P.term_s_hi.push_back(uh);
    // upsilon_i_i (:675-707)
    add_term(3, i, i, np, 0);
    plan_upsilon(ctx, P, fam_items, fixups, KU, P.VU[i], P.VU[i], &Kki, &P.V[i], i, i, false, lo[i], hi[i], diff,
                 &ul, &uh);
    P.term_s_lo.push_back(ul); P.term_s_hi.push_back(uh);
  }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      if (i == j) continue;
      // upsilon_i_j (:640-673)
      add_term(3, i, j, ctx->kernels[k * d + i].n_param(), ctx->kernels[k * d + j].n_param());
      plan_upsilon(ctx, P, fam_items, fixups, KU, P.VU[i], P.VU[j], nullptr, nullptr, i, j, false, lo[i], hi[i],
                   diff, &ul, &uh);
      P.term_s_lo.push_back(ul); P.term_s_hi.push_back(uh);
    }
  P.out_len = out;
  // concatenate the families and assign work units
  int base = 0;
  int64_t wu = 0;
  for (int f = 0; f < 8; ++f) {
    P.fam_first[f] = base;
    P.fam_count[f] = (int)fam_items[f].size();
    P.fam_wu_first[f] = wu;
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
      P.items.push_back(it);
    }
    P.fam_wu_count[f] = wu - P.fam_wu_first[f];
    base += P.fam_count[f];
  }
  P.n_wu = wu;
  for (auto& fx : fixups)
    P.slot_item[fx.first] = fx.second.idx < 0 ? -1 : P.fam_first[fx.second.fam] + fx.second.idx;
  return ASLSD_OK;
}

template <int FAMILY>
static void launch_pair(aslsd_ctx* ctx, const ExactPlan& P, const ItemDev* d_items) {
  if (P.fam_wu_count[FAMILY] == 0) return;
  pair_kernel<FAMILY><<<(unsigned)P.fam_wu_count[FAMILY], PAIR_NT, 0, ctx->stream>>>(
      (const DimDev*)ctx->d_dims, d_items + P.fam_first[FAMILY], P.fam_count[FAMILY], ctx->T_f,
      ctx->d_partials, ctx->d_counts, P.fam_wu_first[FAMILY]);
  ctx->last_launches++;
}
template <int FAMILY>
static void launch_single(aslsd_ctx* ctx, const ExactPlan& P, const ItemDev* d_items) {
  if (P.fam_wu_count[FAMILY] == 0) return;
  single_kernel<FAMILY><<<(unsigned)P.fam_wu_count[FAMILY], PAIR_NT, 0, ctx->stream>>>(
      (const DimDev*)ctx->d_dims, d_items + P.fam_first[FAMILY], P.fam_count[FAMILY], ctx->T_f,
      ctx->d_partials, ctx->d_counts, P.fam_wu_first[FAMILY]);
  ctx->last_launches++;
}

// Launch everything; leaves raw[item][comp] in raw_device.
static int run_plan(aslsd_ctx* ctx, const ExactPlan& P, double* raw_device, unsigned long long* d_item_counts) {
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t n_items = P.items.size();
  if (n_items == 0) return ASLSD_OK;
  int rc;
  if ((rc = ensure_device_buffer(ctx, &ctx->d_items, &ctx->items_cap, n_items * sizeof(ItemDev)))) return rc;
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_partials, &ctx->partials_cap,
                                 (size_t)P.n_wu * ASLSD_MAX_COMP * sizeof(double)))) return rc;
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_counts, &ctx->counts_cap,
                                 ((size_t)P.n_wu + n_items) * sizeof(unsigned long long)))) return rc;
  if (P.n_wu >= (int64_t)2147483647) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: too many work units");
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_items, P.items.data(), n_items * sizeof(ItemDev),
                                cudaMemcpyHostToDevice, ctx->stream));
  const ItemDev* d_items = (const ItemDev*)ctx->d_items;
  ctx->last_launches = 0;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  launch_pair<ITEM_UPS_GAUSS>(ctx, P, d_items);
  launch_pair<ITEM_UPS_EXP>(ctx, P, d_items);
  launch_pair<ITEM_PHI_GAUSS>(ctx, P, d_items);
  launch_pair<ITEM_PHI_EXP>(ctx, P, d_items);
  launch_single<ITEM_UPS0_GAUSS>(ctx, P, d_items);
  launch_single<ITEM_UPS0_EXP>(ctx, P, d_items);
  launch_single<ITEM_PSI_GAUSS>(ctx, P, d_items);
  launch_single<ITEM_PSI_EXP>(ctx, P, d_items);
  const int n_fin = (int)n_items * (ASLSD_MAX_COMP + 1);
  finalize_kernel<<<(n_fin + 127) / 128, 128, 0, ctx->stream>>>(d_items, (int)n_items, ctx->d_partials,
                                                                ctx->d_counts, raw_device, d_item_counts);
  ctx->last_launches++;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timing_pending = true;
  ctx->last_n_wu = P.n_wu;
  ctx->last_n_items = (int)n_items;
  for (int f = 0; f < 8; ++f) { ctx->last_fam_first[f] = P.fam_first[f]; ctx->last_fam_count[f] = P.fam_count[f]; }
  CUDA_TRY(ctx, cudaGetLastError());
  return ASLSD_OK;
}

static int count_pair_terms(aslsd_ctx* ctx, const ExactPlan& P, int64_t* n_pair_terms) {
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

// ---- assembly: raw components -> packed per-term sums ---------------------------------------
static void assemble(const aslsd_ctx* ctx, const ExactPlan& P, const double* x_k, const double* raw,
                     double* out) {
  const int d = ctx->d;
  const int k = P.k;
  const bool diff = (P.flags & ASLSD_DIFF) != 0;
  const double mu_k = x_k[0];
  const KernelHost& KU = ctx->kernels[d * d - 1];
  auto R = [&](int off, int slot, int c) -> double {
    int idx = P.slot_item[off + slot];
    return idx < 0 ? 0. : raw[(size_t)idx * ASLSD_MAX_COMP + c];
  };
  for (size_t t = 0; t < P.terms.size(); ++t) {
    const TermPlan& T = P.terms[t];
    double* o = out + T.out_off;
    for (int c = 0; c < 1 + T.n1 + T.n2; ++c) o[c] = 0.;
    const int voff = P.term_val_off[t], doff = P.term_der_off[t];
    if (T.kind == 0 || T.kind == 2) {
      // psi_i / phi_i: one item per basis of K[k][i]
      const KernelHost& K = ctx->kernels[k * d + T.i];
      const std::vector<Vars>& V = P.V[T.i];
      double S = 0.;
      int slot = 0;
      for (size_t r = 0; r < K.bases.size(); ++r) {
        const BasisHost& b = K.bases[r];
        const bool is_exp = b.type == ASLSD_BASIS_EXPONENTIAL;
        const double w = V[r][0], be = V[r][1];
        const double c0 = R(voff, (int)r, 0), c1 = R(voff, (int)r, 1), c2 = R(voff, (int)r, 2);
        double val, d_w, d_b, d_d = 0.;
        if (T.kind == 2) {
          if (is_exp) { val = w * be * c0; d_w = be * c0; d_b = w * c1; }
          else { val = w * c0; d_w = c0; d_b = (w / be) * c1; d_d = w * c2; }
        } else {
          if (is_exp) { val = w * c0; d_w = c0; d_b = w * c1; }
          else { val = w * c0; d_w = c0; d_b = (w / (be * be)) * c1; d_d = -(w / be) * c2; }
        }
        S += val;
        const double scale = (T.kind == 0) ? mu_k : 1.;
        for (int v = 0; v < b.n_vars(); ++v) {
          if (!((b.free_mask >> v) & 1)) continue;
          if (diff) o[1 + slot] = scale * (v == 0 ? d_w : (v == 1 ? d_b : d_d));
          ++slot;
        }
      }
      if (T.kind == 0) {
        o[0] = mu_k * S;
        o[1 + T.n1] = S;   // derivative with respect to mu_k (ix_func 2)
      } else {
        o[0] = S;
      }
      continue;
    }
    const int Rn = (int)KU.bases.size();
    const bool is_exp = KU.bases[0].type == ASLSD_BASIS_EXPONENTIAL;
    const int cB1 = C_B1, cD1 = C_D1, cB2 = C_B2, cD2 = C_D2;
    (void)is_exp;
    if (T.kind == 3 && T.i != T.j) {
      const std::vector<Vars>&V1 = P.VU[T.i], &V2 = P.VU[T.j];
      double f = 0.;
      for (int p = 0; p < Rn; ++p)
        for (int q = 0; q < Rn; ++q) f += V1[p][0] * V2[q][0] * R(voff, p * Rn + q, C_G);
      o[0] = f;
      if (!diff) continue;
      int slot = 0;
      for (int p = 0; p < Rn; ++p)
        for (int v = 0; v < KU.bases[p].n_vars(); ++v) {
          if (!((KU.bases[p].free_mask >> v) & 1)) continue;
          double s = 0.;
          const int c = v == 0 ? C_G : (v == 1 ? cB1 : cD1);
          for (int q = 0; q < Rn; ++q) s += V2[q][0] * R(voff, p * Rn + q, c);
          if (v != 0) s *= V1[p][0];
          if (slot < T.n1) o[1 + slot] = s;
          ++slot;
        }
      slot = 0;
      for (int q = 0; q < Rn; ++q)
        for (int v = 0; v < KU.bases[q].n_vars(); ++v) {
          if (!((KU.bases[q].free_mask >> v) & 1)) continue;
          double s = 0.;
          const int c = v == 0 ? C_G : (v == 1 ? cB2 : cD2);
          for (int p = 0; p < Rn; ++p) s += V1[p][0] * R(voff, p * Rn + q, c);
          if (v != 0) s *= V2[q][0];
          if (slot < T.n2) o[1 + T.n1 + slot] = s;
          ++slot;
        }
      continue;
    }
    // same-type upsilon (kind 3, i == j) and upsilon at lag zero (kind 1)
    {
      const std::vector<Vars>& VA = P.VU[T.i];
      const std::vector<Vars>& VB = P.V[T.i];
      const KernelHost& KB = ctx->kernels[k * d + T.i];
      double f = 0.;
      for (int p = 0; p < Rn; ++p)
        for (int q = 0; q < Rn; ++q) f += VA[p][0] * VA[q][0] * R(voff, p * Rn + q, C_G);
      o[0] = f;
      if (!diff) continue;
      int slot = 0;
      for (int b = 0; b < Rn; ++b)
        for (int v = 0; v < KB.bases[b].n_vars(); ++v) {
          if (!((KB.bases[b].free_mask >> v) & 1)) continue;
          const double wb = VB[b][0];
          double s;
          if (v == 0) s = wb * R(doff, b * Rn + b, C_SW);
          else s = wb * wb * R(doff, b * Rn + b, v == 1 ? C_SB : C_SD);
          // cross terms with every other basis in both slots (kernel.py:282-299)
          for (int q = 0; q < Rn; ++q) {
            if (q == b) continue;
            if (v == 0) s += VB[q][0] * R(doff, b * Rn + q, C_G);
            else s += wb * VB[q][0] * R(doff, b * Rn + q, v == 1 ? cB1 : cD1);
          }
          for (int p = 0; p < Rn; ++p) {
            if (p == b) continue;
            if (v == 0) s += VB[p][0] * R(doff, p * Rn + b, C_G);
            else s += VB[p][0] * wb * R(doff, p * Rn + b, v == 1 ? cB2 : cD2);
          }
          o[1 + slot] = s;
          ++slot;
        }
    }
  }
}

// ---- C ABI --------------------------------------------------------------------------------------
extern "C" int aslsd_exact_out_len(const aslsd_ctx* ctx, int k, int64_t* out_len, int64_t* n_terms) {
  if (!ctx || !ctx->has_model || k < 0 || k >= ctx->d) return ASLSD_ERR_ARG;
  const int d = ctx->d;
  int64_t out = 0;
  for (int i = 0; i < d; ++i) {
    const int np = ctx->kernels[k * d + i].n_param();
    out += (1 + np + 1) + 3 * (1 + np);
  }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j)
      if (i != j) out += 1 + ctx->kernels[k * d + i].n_param() + ctx->kernels[k * d + j].n_param();
  if (out_len) *out_len = out;
  if (n_terms) *n_terms = 4 * d + d * (d - 1);
  return ASLSD_OK;
}

extern "C" int aslsd_exact_raw_len(aslsd_ctx* ctx, int k, int64_t* raw_len) {
  if (!ctx || !ctx->has_model || k < 0 || k >= ctx->d) return ASLSD_ERR_ARG;
  // upper bound independent of the parameters: every term's full item table
  const int d = ctx->d;
  const int64_t R = (int64_t)ctx->kernels[d * d - 1].bases.size();
  int64_t items = 0;
  for (int i = 0; i < d; ++i) items += 2 * R + 2 * 2 * R * R;
  items += (int64_t)d * (d - 1) * R * R;
  if (raw_len) *raw_len = items * ASLSD_MAX_COMP;
  return ASLSD_OK;
}

extern "C" int aslsd_exact_raw_device(aslsd_ctx* ctx, int k, const double* x_k, int flags, const int64_t* m_lo,
                                      const int64_t* m_hi, double* raw_device, int64_t* n_pair_terms) {
  if (!ctx || !x_k || !raw_device) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_raw_device: bad arguments");
  ExactPlan P;
  int rc = build_plan(ctx, k, x_k, flags, m_lo, m_hi, P);
  if (rc) return rc;
  int64_t raw_len = 0;
  aslsd_exact_raw_len(ctx, k, &raw_len);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, cudaMemsetAsync(raw_device, 0, raw_len * sizeof(double), ctx->stream));
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_counts, &ctx->counts_cap,
                                 ((size_t)P.n_wu + P.items.size() + 1) * sizeof(unsigned long long)))) return rc;
  rc = run_plan(ctx, P, raw_device, ctx->d_counts + P.n_wu);
  if (rc) return rc;
  if (n_pair_terms) rc = count_pair_terms(ctx, P, n_pair_terms);
  return rc;
}

extern "C" int aslsd_exact_assemble(aslsd_ctx* ctx, int k, const double* x_k, int flags, const double* raw_host,
                                    double* out) {
  if (!ctx || !x_k || !raw_host || !out) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_assemble: bad arguments");
  ExactPlan P;
  // the item table (hence the raw layout) depends only on the model, the flags
  // and the parameters, not on the target block
  int rc = build_plan(ctx, k, x_k, flags, nullptr, nullptr, P);
  if (rc) return rc;
  assemble(ctx, P, x_k, raw_host, out);
  return ASLSD_OK;
}

extern "C" int aslsd_exact_sums(aslsd_ctx* ctx, int k, const double* x_k, int flags, const int64_t* m_lo,
                                const int64_t* m_hi, double* out, int64_t* n_pair_terms) {
  if (!ctx || !x_k || !out) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_sums: bad arguments");
  ExactPlan P;
  int rc = build_plan(ctx, k, x_k, flags, m_lo, m_hi, P);
  if (rc) return rc;
  const size_t n_items = P.items.size();
  const size_t raw_bytes = std::max<size_t>(1, n_items) * ASLSD_MAX_COMP * sizeof(double);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_raw, &ctx->raw_cap, raw_bytes))) return rc;
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_counts, &ctx->counts_cap,
                                 ((size_t)P.n_wu + n_items + 1) * sizeof(unsigned long long)))) return rc;
  if ((rc = ensure_pinned_buffer(ctx, raw_bytes))) return rc;
  if ((rc = run_plan(ctx, P, ctx->d_raw, ctx->d_counts + P.n_wu))) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_raw, raw_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  assemble(ctx, P, x_k, (const double*)ctx->h_pinned, out);
  if (n_pair_terms) rc = count_pair_terms(ctx, P, n_pair_terms);
  return rc;
}

// Number of closed-form evaluations of the last exact-sum call: basis-pair
// evaluations of the pair kernels (upsilon, phi) and of the single kernels.
extern "C" int aslsd_last_eval_counts(aslsd_ctx* ctx, int64_t* pair_evals, int64_t* single_evals) {
  if (!ctx) return ASLSD_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  std::vector<unsigned long long> h(std::max(1, ctx->last_n_items));
  if (ctx->last_n_items > 0) {
    CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), ctx->d_counts + ctx->last_n_wu,
                                  sizeof(unsigned long long) * ctx->last_n_items, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  int64_t pe = 0, se = 0;
  for (int f = 0; f < 8; ++f) {
    const bool pair = (f == ITEM_UPS_EXP || f == ITEM_UPS_GAUSS || f == ITEM_PHI_EXP || f == ITEM_PHI_GAUSS);
    for (int x = 0; x < ctx->last_fam_count[f]; ++x)
      (pair ? pe : se) += (int64_t)h[ctx->last_fam_first[f] + x];
  }
  if (pair_evals) *pair_evals = pe;
  if (single_evals) *single_evals = se;
  return ASLSD_OK;
}
