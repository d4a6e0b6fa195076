// Host-side planning shared by the exact and the sampled sums of one sub-problem k:
// named terms -> work items (one basis / one basis pair each) and the linear map
// from the items' raw components to the packed per-term outputs.
//
// Mirrors the composition rules of the reference:
//   KernelModel.make_*            functionals/kernels/kernel.py:224-332
//   BasisKernel.make_vars         functionals/kernels/basis_kernel.py:364-401
//   MHP.make_functionals          models/hawkes/linear/mhp.py:475-540 (late binding)
//   term order / output shapes    optimize/estimators/mhp_exact_estimator.py:141-170, :709-761
#pragma once

#include <array>
#include <cmath>

#include "common.cuh"

typedef std::array<double, 3> Vars;

struct SumPlan {
  int k = -1, flags = 0;
  std::vector<TermPlan> terms;
  std::vector<ItemDev> items;       // grouped by family (exact) or by term (sampled)
  int fam_first[ITEM_N_FAMILIES], fam_count[ITEM_N_FAMILIES];
  int64_t fam_wu_first[ITEM_N_FAMILIES], fam_wu_count[ITEM_N_FAMILIES];
  int64_t n_wu = 0;
  // per-term tables of item indices (-1: empty window, contributes zero)
  std::vector<int> slot_item;
  std::vector<int> term_val_off, term_der_off;  // offsets into slot_item
  int64_t out_len = 0;
  std::vector<std::vector<Vars>> V;   // per source dim i: vars of K[k][i] with x_ki
  std::vector<std::vector<Vars>> VU;  // per source dim i: vars of KU with x_ki
  std::vector<double> term_s_lo, term_s_hi;  // union window per term (pair counting)
};

// Terms in the reference's evaluation order: psi_i, upsilonzero_i, phi_i,
// upsilon_i_i for i = 0..d-1, then upsilon_i_j (i != j) in itertools.product order.
//   by_term = false: items grouped by family with lag windows and work units (exact sums)
//   by_term = true : items grouped by term, no window pruning (sampled sums)
int build_plan(aslsd_ctx* ctx, int k, const double* x_k, int flags, const int64_t* m_lo,
               const int64_t* m_hi, bool by_term, SumPlan& P);

// Re-derive the per-basis variables (P.V, P.VU) of an existing plan from a new x_k.
int plan_vars(aslsd_ctx* ctx, int k, const double* x_k, SumPlan& P);
// Append the non-omega variables of x_k to `key` (the work items depend on x_k only
// through them).
void plan_shape_key(const aslsd_ctx* ctx, int k, const double* x_k, std::vector<double>& key);

// Index of a named term in SumPlan::terms (-1 if absent).
int find_term(const aslsd_ctx* ctx, int kind, int i, int j);

// Symbolic omega factor: coefficient = ((w[a] x w[b]) x g) / gd, with a, b indices into a
// per-sub-problem table of the omega values (0 = the constant one) and g, gd functions of the
// non-omega variables only (the divisor is kept apart so that w / beta rounds as in emit_term<double>).  emit_term_w instantiated with it yields the structure of the linear map once per plan;
// the device assembly (exact.cu) then evaluates the coefficients from the table of every call.
struct SymW {
  int32_t a = 0, b = 0;
  double g = 1., gd = 1.;
  SymW() {}
  SymW(double v) : g(v) {}
  SymW(int32_t ix, double v) : a(ix), g(v) {}
};
inline SymW operator*(const SymW& x, const SymW& y) {
  SymW r;
  r.g = x.g * y.g;
  r.gd = x.gd * y.gd;
  const int32_t idx[4] = {x.a, x.b, y.a, y.b};
  int n = 0;
  for (int q = 0; q < 4; ++q)
    if (idx[q] != 0) {
      if (n == 0) r.a = idx[q];
      else if (n == 1) r.b = idx[q];
      else r.g = std::nan("");   // more than two omega factors: no such coefficient exists
      ++n;
    }
  return r;
}
inline SymW operator*(const SymW& x, double s) { SymW r = x; r.g = x.g * s; return r; }
inline SymW operator*(double s, const SymW& x) { return x * s; }
inline SymW operator/(const SymW& x, double s) { SymW r = x; r.gd = x.gd * s; return r; }
inline SymW operator-(const SymW& x) { SymW r = x; r.g = -x.g; return r; }

// Linear map raw components -> packed outputs of one term.  The sink is called
// as sink(c, item, comp, coef, outer): output c of the term (0 = function value,
// 1.. = derivatives) receives outer * (coef * raw[item][comp]); item < 0 never
// reaches it.  The two-level product follows the reference's association
// (e.g. psi term: mu_k * (omega * (1 - e^{-beta t})), mhp_stratified_estimator.py
// :1106-1108) so that per-sample values of near-constant strata come out
// bit-identical and numpy's rounding residue in np.std is reproduced.
// W: type of the omega factors (double, or SymW for the structure alone); omV(i, r) / omVU(i, r):
// omega of basis r of K[k][i] read with the structure of K[k][i] / of KU; the other variables
// come from P.V / P.VU.
template <class W, class OmV, class OmVU, class Sink>
void emit_term_w(const aslsd_ctx* ctx, const SumPlan& P, size_t t, W mu_k, OmV&& omV, OmVU&& omVU, Sink&& sink) {
  const int d = ctx->d;
  const int k = P.k;
  const bool diff = (P.flags & ASLSD_DIFF) != 0;
  const KernelHost& KU = ctx->kernels[d * d - 1];
  const TermPlan& T = P.terms[t];
  const int voff = P.term_val_off[t], doff = P.term_der_off[t];
  const W one(1.);
  auto put = [&](int c, int off, int slot, int comp, W coef, W outer) {
    const int idx = P.slot_item[off + slot];
    if (idx >= 0) sink(c, idx, comp, coef, outer);
  };
  // factor of every value component (C_G) of the basis pair (p, q), by the types of the two bases:
  // exp x exp beta1 beta2 / (beta1 + beta2), exp x gauss beta1, gauss x exp beta2, gauss x gauss 1
  auto is_exp = [&](int r) { return KU.bases[r].type == ASLSD_BASIS_EXPONENTIAL; };
  auto prat = [&](int p, int q, const Vars& a, const Vars& b) {
    if (is_exp(p)) return is_exp(q) ? (a[1] * b[1]) / (a[1] + b[1]) : a[1];
    return is_exp(q) ? b[1] : 1.;
  };
  // component read by the derivative in omega of basis `side` (1 or 2) of the pair: mixed pairs
  // keep the value under the reference's overflow guard in slot C_SW (basis_math.cuh); the
  // gauss x exp form does not guard its omega_2 derivative (basis_kernel_gaussian.py:196-198)
  auto omega_comp = [&](int p, int q, int side) {
    if (is_exp(p) == is_exp(q)) return (int)C_G;
    if (!is_exp(p) && side == 2) return (int)C_G;
    return (int)C_SW;
  };
  if (T.kind == 0 || T.kind == 2) {
    // psi_i / phi_i: one item per basis of K[k][i]
    const KernelHost& K = ctx->kernels[k * d + T.i];
    const std::vector<Vars>& V = P.V[T.i];
    const W scale = (T.kind == 0) ? mu_k : one;
    int slot = 0;
    for (size_t r = 0; r < K.bases.size(); ++r) {
      const BasisHost& b = K.bases[r];
      const bool is_exp = b.type == ASLSD_BASIS_EXPONENTIAL;
      const W w = omV(T.i, (int)r);
      const double be = V[r][1];
      W cv, cw, cb, cd(0.);   // coefficients of comp 0 (value), comp 0 (d omega), comp 1, comp 2
      if (T.kind == 2) {
        if (is_exp) { cv = w * be; cw = W(be); cb = w; }
        else { cv = w; cw = one; cb = w / be; cd = w; }
      } else {
        if (is_exp) { cv = w; cw = one; cb = w; }
        else { cv = w; cw = one; cb = w / (be * be); cd = -(w / be); }
      }
      put(0, voff, (int)r, 0, cv, scale);
      if (T.kind == 0) put(1 + T.n1, voff, (int)r, 0, cv, one);   // d/d mu_k (ix_func 2)
      for (int v = 0; v < b.n_vars(); ++v) {
        if (!((b.free_mask >> v) & 1)) continue;
        if (diff) put(1 + slot, voff, (int)r, v, v == 0 ? cw : (v == 1 ? cb : cd), scale);
        ++slot;
      }
    }
    return;
  }
  const int Rn = (int)KU.bases.size();
  if (T.kind == 3 && T.i != T.j) {
    const std::vector<Vars>&V1 = P.VU[T.i], &V2 = P.VU[T.j];
    for (int p = 0; p < Rn; ++p)
      for (int q = 0; q < Rn; ++q)
        put(0, voff, p * Rn + q, C_G, omVU(T.i, p) * omVU(T.j, q) * prat(p, q, V1[p], V2[q]), one);
    if (!diff) return;
    int slot = 0;
    for (int p = 0; p < Rn; ++p)
      for (int v = 0; v < KU.bases[p].n_vars(); ++v) {
        if (!((KU.bases[p].free_mask >> v) & 1)) continue;
        if (slot < T.n1)
          for (int q = 0; q < Rn; ++q)
            put(1 + slot, voff, p * Rn + q, v == 0 ? omega_comp(p, q, 1) : (v == 1 ? C_B1 : C_D1),
                v == 0 ? omVU(T.j, q) * prat(p, q, V1[p], V2[q]) : omVU(T.i, p) * omVU(T.j, q), one);
        ++slot;
      }
    slot = 0;
    for (int q = 0; q < Rn; ++q)
      for (int v = 0; v < KU.bases[q].n_vars(); ++v) {
        if (!((KU.bases[q].free_mask >> v) & 1)) continue;
        if (slot < T.n2)
          for (int p = 0; p < Rn; ++p)
            put(1 + T.n1 + slot, voff, p * Rn + q, v == 0 ? omega_comp(p, q, 2) : (v == 1 ? C_B2 : C_D2),
                v == 0 ? omVU(T.i, p) * prat(p, q, V1[p], V2[q]) : omVU(T.i, p) * omVU(T.j, q), one);
        ++slot;
      }
    return;
  }
  // same-type upsilon (kind 3, i == j) and upsilon at lag zero (kind 1)
  const std::vector<Vars>& VA = P.VU[T.i];
  const std::vector<Vars>& VB = P.V[T.i];
  const KernelHost& KB = ctx->kernels[k * d + T.i];
  for (int p = 0; p < Rn; ++p)
    for (int q = 0; q < Rn; ++q)
      put(0, voff, p * Rn + q, C_G, omVU(T.i, p) * omVU(T.i, q) * prat(p, q, VA[p], VA[q]), one);
  if (!diff) return;
  int slot = 0;
  for (int b = 0; b < Rn; ++b)
    for (int v = 0; v < KB.bases[b].n_vars(); ++v) {
      if (!((KB.bases[b].free_mask >> v) & 1)) continue;
      const W wb = omV(T.i, b);
      // own-basis closed form, then the cross terms with every other basis in
      // both slots (kernel.py:282-299)
      if (v == 0) put(1 + slot, doff, b * Rn + b, C_SW, is_exp(b) ? wb * VB[b][1] : wb, one);
      else put(1 + slot, doff, b * Rn + b, v == 1 ? C_SB : C_SD, wb * wb, one);
      for (int q = 0; q < Rn; ++q) {
        if (q == b) continue;
        if (v == 0) put(1 + slot, doff, b * Rn + q, omega_comp(b, q, 1), omV(T.i, q) * prat(b, q, VB[b], VB[q]), one);
        else put(1 + slot, doff, b * Rn + q, v == 1 ? C_B1 : C_D1, wb * omV(T.i, q), one);
      }
      for (int p = 0; p < Rn; ++p) {
        if (p == b) continue;
        if (v == 0) put(1 + slot, doff, p * Rn + b, omega_comp(p, b, 2), omV(T.i, p) * prat(p, b, VB[p], VB[b]), one);
        else put(1 + slot, doff, p * Rn + b, v == 1 ? C_B2 : C_D2, omV(T.i, p) * wb, one);
      }
      ++slot;
    }
}

template <class Sink>
void emit_term(const aslsd_ctx* ctx, const SumPlan& P, size_t t, const double* x_k, Sink&& sink) {
  emit_term_w<double>(ctx, P, t, x_k[0], [&](int i, int r) { return P.V[i][r][0]; },
                      [&](int i, int r) { return P.VU[i][r][0]; }, sink);
}
