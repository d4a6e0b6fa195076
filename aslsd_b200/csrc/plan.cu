// Planning of the LSE sums of one sub-problem k: terms -> work items.
// Replaces the term loop of MHPExactEstim.compute_objective
// (optimize/estimators/mhp_exact_estimator.py:709-761) and the closure wiring of
// MHP.make_functionals (models/hawkes/linear/mhp.py:475-540).
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "pair_tiles.h"


// Effective support of the kernels: lags outside contribute < 1e-17 of a term
// (agreement with the untruncated sum ~1e-15).
static const double Z_GAUSS = 9.0;     // Gaussian half-width in standard deviations
static const double C_EXP = 42.0;      // exponential support in units of 1/beta
static const double X_INTERIOR = 10.0; // upper integration limit treated as +inf beyond this
static const double C_INTERIOR = 60.0; // exp(-x) treated as 0 beyond this


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

static void ups_window(int type1, int type2, const Vars& v1, const Vars& v2, double* s_lo, double* s_hi,
                       double* t_int) {
  const bool e1 = type1 == ASLSD_BASIS_EXPONENTIAL, e2 = type2 == ASLSD_BASIS_EXPONENTIAL;
  if (e1 && e2) {
    *s_lo = 0.;
    *s_hi = C_EXP / v2[1];
    *t_int = C_INTERIOR / (v1[1] + v2[1]);
  } else if (!e1 && !e2) {
    const double sig = std::sqrt(v1[1] * v1[1] + v2[1] * v2[1]);
    const double dd = v2[2] - v1[2];
    *s_lo = std::max(0., dd - Z_GAUSS * sig);
    *s_hi = dd + Z_GAUSS * sig;
    const double bs = v1[1] * v2[1] / sig;
    const double r1 = v2[1] * v2[1] / (sig * sig), r2 = v1[1] * v1[1] / (sig * sig);
    *t_int = X_INTERIOR * bs + r1 * v1[2] + r2 * std::max(v2[2], 0.);
  } else if (e1) {
    // exponential x Gaussian: the exponential (from 0) convolved with the Gaussian bump at delta2:
    // beyond delta2 + z beta2 the product e^{beta1 s} x (Gaussian tail) is below e^{-z^2 / 2};
    // upper limit (t - delta2 + s + beta1 beta2^2) / beta2 >= X for every s >= 0
    *s_lo = 0.;
    *s_hi = std::max(0., v2[2]) + Z_GAUSS * v2[1];
    *t_int = std::max(0., v2[2]) + X_INTERIOR * v2[1];
  } else {
    // Gaussian x exponential: e^{-beta2 s} times a factor that depends on s only through the
    // limits of the (verbatim) beta1 / delta1 branches, (t - delta1 - s + beta2 beta1^2) / beta1
    *s_lo = 0.;
    *s_hi = C_EXP / v2[1];
    *t_int = std::max(0., v1[2]) + *s_hi + X_INTERIOR * v1[1];
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

// Per-event weights (marked models) and activity of the term being planned.
struct TermWeights {
  int64_t wa = -1, wb = -1;   // offsets into the context's weight block
  bool active = true;         // false: the term is masked out, its slots stay empty
};

// Build the work items of an upsilon-type term.
//   VA1/VA2: variables of the value closure (structure KU), VB: variables of the
//   derivative closure of a same-type term (K[k][i]); merged when equal.
static void plan_upsilon(const aslsd_ctx* ctx, SumPlan& P, std::vector<ItemDev> fam_items[ITEM_N_FAMILIES],
                         std::vector<std::pair<int, ItemRef>>& fixups, const KernelHost& KU,
                         const std::vector<Vars>& VA1, const std::vector<Vars>& VA2,
                         const KernelHost* KB, const std::vector<Vars>* VB, int tgt, int src, bool single,
                         int64_t m_lo, int64_t m_hi, bool diff, bool by_term, double* u_lo, double* u_hi,
                         const TermWeights& W) {
  const int R1 = (int)KU.bases.size(), R2 = R1;
  const bool sim = (KB != nullptr);
  const bool merged = sim && vars_equal(VA1, *VB);
  // family of the basis pair (p, q); all kernels share KU's sequence of basis types (aslsd_set_model)
  auto pair_family = [&](int p, int q) {
    const bool e1 = KU.bases[p].type == ASLSD_BASIS_EXPONENTIAL, e2 = KU.bases[q].type == ASLSD_BASIS_EXPONENTIAL;
    if (e1 && e2) return single ? ITEM_UPS0_EXP : ITEM_UPS_EXP;
    if (!e1 && !e2) return single ? ITEM_UPS0_GAUSS : ITEM_UPS_GAUSS;
    if (e1) return single ? ITEM_UPS0_EG : ITEM_UPS_EG;
    return single ? ITEM_UPS0_GE : ITEM_UPS_GE;
  };
  *u_lo = 1e300;
  *u_hi = 0.;
  auto emit = [&](int p, int q, const Vars& v1, const Vars& v2, int mask, bool own, const Vars& vo) {
    if (!W.active) return ItemRef{0, -1};
    ItemDev it;
    std::memset(&it, 0, sizeof(it));
    it.wa_off = W.wa;
    it.wb_off = W.wb;
    const int fam = pair_family(p, q);
    const int sf = by_term ? 0 : fam;   // storage list: one per family, or a single one in term order
    it.family = fam;
    it.tgt_dim = tgt;
    it.src_dim = src;
    it.same = (tgt == src) ? 1 : 0;
    it.m_lo = m_lo;
    it.m_hi = m_hi;
    it.b1 = v1[1]; it.d1 = v1[2]; it.b2 = v2[1]; it.d2 = v2[2];
    it.bo = vo[1]; it.dlo = vo[2];
    ups_window(KU.bases[p].type, KU.bases[q].type, v1, v2, &it.s_lo, &it.s_hi, &it.t_int);
    if (own) {
      double oh, ot;
      own_window(KU.bases[p].type, vo, &oh, &ot);
      it.s_lo = 0.;
      it.s_hi = std::max(it.s_hi, oh);
      it.t_int = std::max(it.t_int, ot);
      if (vo[1] == v1[1] && vo[1] == v2[1] && vo[2] == v1[2] && vo[2] == v2[2]) mask |= MASK_OWN_SAME;
    }
    it.mask = mask;
    if (!by_term && it.s_hi < 0.) return ItemRef{sf, -1};   // empty lag window: contributes exactly zero
    *u_lo = std::min(*u_lo, it.s_lo);
    *u_hi = std::max(*u_hi, it.s_hi);
    fam_items[sf].push_back(it);
    return ItemRef{sf, (int)fam_items[sf].size() - 1};
  };
  // mixed pairs: the omega derivatives read the guarded copy of the value (slot C_SW)
  auto guard_mask = [&](int p, int q, const std::vector<BasisHost>& B) {
    if (KU.bases[p].type == KU.bases[q].type) return 0;
    return ((B[p].free_mask | B[q].free_mask) & 1) ? MASK(C_SW) : 0;
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
        mask |= guard_mask(p, q, BU);
      }
      if (diff && merged) {
        const std::vector<BasisHost>& BB = KB->bases;
        if (p != q) {
          if ((BB[p].free_mask >> 1) & 1) mask |= MASK(C_B1);
          if ((BB[p].free_mask >> 2) & 1) mask |= MASK(C_D1);
          if ((BB[q].free_mask >> 1) & 1) mask |= MASK(C_B2);
          if ((BB[q].free_mask >> 2) & 1) mask |= MASK(C_D2);
          mask |= guard_mask(p, q, BB);
        } else {
          if ((BB[p].free_mask >> 0) & 1) mask |= MASK(C_SW);
          if ((BB[p].free_mask >> 1) & 1) mask |= MASK(C_SB);
          if ((BB[p].free_mask >> 2) & 1) mask |= MASK(C_SD);
          own = (mask & (MASK(C_SW) | MASK(C_SB) | MASK(C_SD))) != 0;
        }
      }
      const ItemRef ref = emit(p, q, VA1[p], VA2[q], mask, own, own ? (*VB)[p] : VA1[p]);
      fixups.push_back({(int)P.slot_item.size(), ref});
      P.slot_item.push_back(ref.idx);
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
          mask |= guard_mask(p, q, BB);
        } else {
          if ((BB[p].free_mask >> 0) & 1) mask |= MASK(C_SW);
          if ((BB[p].free_mask >> 1) & 1) mask |= MASK(C_SB);
          if ((BB[p].free_mask >> 2) & 1) mask |= MASK(C_SD);
          own = (mask & (MASK(C_SW) | MASK(C_SB) | MASK(C_SD))) != 0;
        }
        const ItemRef ref = emit(p, q, (*VB)[p], (*VB)[q], mask, own, (*VB)[p]);
        fixups.push_back({(int)P.slot_item.size(), ref});
        P.slot_item.push_back(ref.idx);
      }
  } else if (sim && merged) {
    P.term_der_off.back() = P.term_val_off.back();
  }
  if (*u_lo > *u_hi) { *u_lo = 0.; *u_hi = -1.; }
}

static void plan_basis_items(SumPlan& P, std::vector<ItemDev> fam_items[ITEM_N_FAMILIES],
                             std::vector<std::pair<int, ItemRef>>& fixups, const KernelHost& K,
                             const std::vector<Vars>& V, bool is_phi, int tgt, int src, int64_t m_lo,
                             int64_t m_hi, bool diff, bool by_term, double* u_lo, double* u_hi,
                             const TermWeights& W) {
  P.term_val_off.push_back((int)P.slot_item.size());
  *u_lo = 1e300;
  *u_hi = 0.;
  for (size_t r = 0; r < K.bases.size(); ++r) {
    const BasisHost& b = K.bases[r];
    const bool is_exp = b.type == ASLSD_BASIS_EXPONENTIAL;
    if (!W.active) {
      fixups.push_back({(int)P.slot_item.size(), ItemRef{0, -1}});
      P.slot_item.push_back(-1);
      continue;
    }
    ItemDev it;
    std::memset(&it, 0, sizeof(it));
    it.wa_off = W.wa;
    it.wb_off = W.wb;
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
    // phi has no T_f - t_m dependence; psi is constant once its upper limit saturates
    if (is_phi) it.t_int = 1e300;
    else it.t_int = is_exp ? C_INTERIOR / V[r][1] : V[r][2] + X_INTERIOR * V[r][1];
    const int sf = by_term ? 0 : it.family;
    fam_items[sf].push_back(it);
    const int idx = (int)fam_items[sf].size() - 1;
    fixups.push_back({(int)P.slot_item.size(), ItemRef{sf, idx}});
    P.slot_item.push_back(idx);
  }
  P.term_der_off.push_back(P.term_val_off.back());
}

// parameter slices of x_k (mhp.py:276-299 interval_map) -> full variables of every basis
int plan_vars(aslsd_ctx* ctx, int k, const double* x_k, SumPlan& P) {
  const int d = ctx->d;
  const KernelHost& KU = ctx->kernels[d * d - 1];
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
  return ASLSD_OK;
}

// The non-omega variables of x_k (beta, delta of every basis, in order): together
// with the model, the flags and the target block they determine the work items.
void plan_shape_key(const aslsd_ctx* ctx, int k, const double* x_k, std::vector<double>& key) {
  const int d = ctx->d;
  int pos = 1;
  for (int i = 0; i < d; ++i)
    for (auto& b : ctx->kernels[k * d + i].bases)
      for (int v = 0; v < b.n_vars(); ++v)
        if ((b.free_mask >> v) & 1) {
          if (v > 0) key.push_back(x_k[pos]);
          ++pos;
        }
}

int build_plan(aslsd_ctx* ctx, int k, const double* x_k, int flags, const int64_t* m_lo,
               const int64_t* m_hi, bool by_term, SumPlan& P) {
  const int d = ctx->d;
  if (!ctx->has_model) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: no model set");
  const bool have_events = ctx->d_times != nullptr && (int)ctx->n_events.size() == d;
  // the work items themselves do not depend on the events: target ranges are patched in by
  // the caller (build_work); only the by-term plans of the sampled sums read them here
  if (!have_events && by_term) return aslsd_fail(ctx, ASLSD_ERR_ARG, "sampled sums: no events set");
  if (k < 0 || k >= d) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact sums: bad k");
  P.k = k;
  P.flags = flags;
  const bool diff = (flags & ASLSD_DIFF) != 0;
  const KernelHost& KU = ctx->kernels[d * d - 1];
  std::vector<int64_t> lo(d), hi(d);
  for (int i = 0; i < d; ++i) {
    lo[i] = m_lo ? std::max<int64_t>(0, m_lo[i]) : 0;
    const int64_t n_i = have_events ? ctx->n_events[i] : 0;
    hi[i] = m_hi ? std::min<int64_t>(n_i, m_hi[i]) : n_i;
  }
  int rc = plan_vars(ctx, k, x_k, P);
  if (rc) return rc;
  std::vector<ItemDev> fam_items[ITEM_N_FAMILIES];
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
  // weights / activity of a term: role 0 phi sources, 1 upsilon targets, 2 upsilon sources,
  // 3 psi targets, 4 upsilonzero targets (aslsd_exact_sums_weighted)
  auto weight_of = [&](int role, int dim) -> int64_t {
    if (!ctx->roles) return -1;
    const int v = ctx->roles[role * d + dim];
    return v < 0 ? -1 : ctx->weight_off[v];
  };
  auto term_w = [&](int64_t wa, int64_t wb) {
    TermWeights W;
    W.wa = wa;
    W.wb = wb;
    W.active = by_term || !ctx->term_mask || ctx->term_mask[P.terms.size() - 1] != 0;
    return W;
  };
  // in term order the single storage list gives every term a contiguous item range
  auto close_term = [&](size_t first) {
    if (!by_term) return;
    P.terms.back().item_first = (int)first;
    P.terms.back().item_count = (int)(fam_items[0].size() - first);
  };
  for (int i = 0; i < d; ++i) {
    const KernelHost& Kki = ctx->kernels[k * d + i];
    const int np = Kki.n_param();
    size_t first;
    // psi_i (:543-578)
    add_term(0, i, i, np, 1);
    first = fam_items[0].size();
    plan_basis_items(P, fam_items, fixups, Kki, P.V[i], false, i, i, lo[i], hi[i], diff, by_term, &ul, &uh,
                     term_w(weight_of(3, i), -1));
    close_term(first);
    P.term_s_lo.push_back(0.); P.term_s_hi.push_back(-1.);
    // upsilonzero_i (:580-609)
    add_term(1, i, i, np, 0);
    first = fam_items[0].size();
    plan_upsilon(ctx, P, fam_items, fixups, KU, P.VU[i], P.VU[i], &Kki, &P.V[i], i, i, true, lo[i], hi[i], diff,
                 by_term, &ul, &uh, term_w(weight_of(4, i), -1));
    close_term(first);
    P.term_s_lo.push_back(0.); P.term_s_hi.push_back(-1.);
    // phi_i (:611-638): targets of type k, sources of type i
    add_term(2, i, i, np, 0);
    first = fam_items[0].size();
    plan_basis_items(P, fam_items, fixups, Kki, P.V[i], true, k, i, lo[k], hi[k], diff, by_term, &ul, &uh,
                     term_w(-1, weight_of(0, i)));
    close_term(first);
    P.term_s_lo.push_back(ul); P.term_s_hi.push_back(uh);
    // upsilon_i_i (:675-707)
    add_term(3, i, i, np, 0);
    first = fam_items[0].size();
    plan_upsilon(ctx, P, fam_items, fixups, KU, P.VU[i], P.VU[i], &Kki, &P.V[i], i, i, false, lo[i], hi[i], diff,
                 by_term, &ul, &uh, term_w(weight_of(1, i), weight_of(2, i)));
    close_term(first);
    P.term_s_lo.push_back(ul); P.term_s_hi.push_back(uh);
  }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      if (i == j) continue;
      // upsilon_i_j (:640-673)
      add_term(3, i, j, ctx->kernels[k * d + i].n_param(), ctx->kernels[k * d + j].n_param());
      const size_t first = fam_items[0].size();
      plan_upsilon(ctx, P, fam_items, fixups, KU, P.VU[i], P.VU[j], nullptr, nullptr, i, j, false, lo[i], hi[i],
                   diff, by_term, &ul, &uh, term_w(weight_of(1, i), weight_of(2, j)));
      close_term(first);
      P.term_s_lo.push_back(ul); P.term_s_hi.push_back(uh);
    }
  P.out_len = out;
  if (by_term) {
    P.items.swap(fam_items[0]);
    for (int f = 0; f < ITEM_N_FAMILIES; ++f) { P.fam_first[f] = 0; P.fam_count[f] = 0; P.fam_wu_first[f] = 0; P.fam_wu_count[f] = 0; }
    P.n_wu = 0;
    for (auto& fx : fixups) P.slot_item[fx.first] = fx.second.idx;
    return ASLSD_OK;
  }
  // concatenate the families and assign work units
  int base = 0;
  int64_t wu = 0;
  for (int f = 0; f < ITEM_N_FAMILIES; ++f) {
    P.fam_first[f] = base;
    P.fam_count[f] = (int)fam_items[f].size();
    P.fam_wu_first[f] = wu;
    const bool pair = item_pair_family(f);
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


// terms are laid out psi_i, upsilonzero_i, phi_i, upsilon_i_i for each i, then the
// d(d-1) cross terms in itertools.product order
int find_term(const aslsd_ctx* ctx, int kind, int i, int j) {
  const int d = ctx->d;
  if (i < 0 || i >= d) return -1;
  if (kind >= 0 && kind <= 2) return 4 * i + kind;
  if (kind != 3 || j < 0 || j >= d) return -1;
  if (i == j) return 4 * i + 3;
  return 4 * d + i * (d - 1) + (j < i ? j : j - 1);
}
