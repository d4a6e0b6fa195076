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

#include "common.cuh"

typedef std::array<double, 3> Vars;

struct SumPlan {
  int k = -1, flags = 0;
  std::vector<TermPlan> terms;
  std::vector<ItemDev> items;       // grouped by family (exact) or by term (sampled)
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

// Linear map raw components -> packed outputs of one term.  The sink is called
// as sink(c, item, comp, coef, outer): output c of the term (0 = function value,
// 1.. = derivatives) receives outer * (coef * raw[item][comp]); item < 0 never
// reaches it.  The two-level product follows the reference's association
// (e.g. psi term: mu_k * (omega * (1 - e^{-beta t})), mhp_stratified_estimator.py
// :1106-1108) so that per-sample values of near-constant strata come out
// bit-identical and numpy's rounding residue in np.std is reproduced.
template <class Sink>
void emit_term(const aslsd_ctx* ctx, const SumPlan& P, size_t t, const double* x_k, Sink&& sink) {
  const int d = ctx->d;
  const int k = P.k;
  const bool diff = (P.flags & ASLSD_DIFF) != 0;
  const double mu_k = x_k[0];
  const KernelHost& KU = ctx->kernels[d * d - 1];
  const TermPlan& T = P.terms[t];
  const int voff = P.term_val_off[t], doff = P.term_der_off[t];
  auto put = [&](int c, int off, int slot, int comp, double coef, double outer = 1.) {
    const int idx = P.slot_item[off + slot];
    if (idx >= 0) sink(c, idx, comp, coef, outer);
  };
  const bool ups_exp = KU.bases[0].type == ASLSD_BASIS_EXPONENTIAL;
  // exponential upsilon: beta1 beta2 / (beta1 + beta2) multiplies every C_G component
  auto prat = [&](const Vars& a, const Vars& b) { return ups_exp ? (a[1] * b[1]) / (a[1] + b[1]) : 1.; };
  if (T.kind == 0 || T.kind == 2) {
    // psi_i / phi_i: one item per basis of K[k][i]
    const KernelHost& K = ctx->kernels[k * d + T.i];
    const std::vector<Vars>& V = P.V[T.i];
    const double scale = (T.kind == 0) ? mu_k : 1.;
    int slot = 0;
    for (size_t r = 0; r < K.bases.size(); ++r) {
      const BasisHost& b = K.bases[r];
      const bool is_exp = b.type == ASLSD_BASIS_EXPONENTIAL;
      const double w = V[r][0], be = V[r][1];
      double cv, cw, cb, cd = 0.;   // coefficients of comp 0 (value), comp 0 (d omega), comp 1, comp 2
      if (T.kind == 2) {
        if (is_exp) { cv = w * be; cw = be; cb = w; }
        else { cv = w; cw = 1.; cb = w / be; cd = w; }
      } else {
        if (is_exp) { cv = w; cw = 1.; cb = w; }
        else { cv = w; cw = 1.; cb = w / (be * be); cd = -(w / be); }
      }
      put(0, voff, (int)r, 0, cv, scale);
      if (T.kind == 0) put(1 + T.n1, voff, (int)r, 0, cv);   // d/d mu_k (ix_func 2)
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
        put(0, voff, p * Rn + q, C_G, V1[p][0] * V2[q][0] * prat(V1[p], V2[q]));
    if (!diff) return;
    int slot = 0;
    for (int p = 0; p < Rn; ++p)
      for (int v = 0; v < KU.bases[p].n_vars(); ++v) {
        if (!((KU.bases[p].free_mask >> v) & 1)) continue;
        const int c = v == 0 ? C_G : (v == 1 ? C_B1 : C_D1);
        if (slot < T.n1)
          for (int q = 0; q < Rn; ++q)
            put(1 + slot, voff, p * Rn + q, c,
                v == 0 ? V2[q][0] * prat(V1[p], V2[q]) : V1[p][0] * V2[q][0]);
        ++slot;
      }
    slot = 0;
    for (int q = 0; q < Rn; ++q)
      for (int v = 0; v < KU.bases[q].n_vars(); ++v) {
        if (!((KU.bases[q].free_mask >> v) & 1)) continue;
        const int c = v == 0 ? C_G : (v == 1 ? C_B2 : C_D2);
        if (slot < T.n2)
          for (int p = 0; p < Rn; ++p)
            put(1 + T.n1 + slot, voff, p * Rn + q, c,
                v == 0 ? V1[p][0] * prat(V1[p], V2[q]) : V1[p][0] * V2[q][0]);
        ++slot;
      }
    return;
  }
  // same-type upsilon (kind 3, i == j) and upsilon at lag zero (kind 1)
  const std::vector<Vars>& VA = P.VU[T.i];
  const std::vector<Vars>& VB = P.V[T.i];
  const KernelHost& KB = ctx->kernels[k * d + T.i];
  for (int p = 0; p < Rn; ++p)
    for (int q = 0; q < Rn; ++q) put(0, voff, p * Rn + q, C_G, VA[p][0] * VA[q][0] * prat(VA[p], VA[q]));
  if (!diff) return;
  int slot = 0;
  for (int b = 0; b < Rn; ++b)
    for (int v = 0; v < KB.bases[b].n_vars(); ++v) {
      if (!((KB.bases[b].free_mask >> v) & 1)) continue;
      const double wb = VB[b][0];
      // own-basis closed form, then the cross terms with every other basis in
      // both slots (kernel.py:282-299)
      if (v == 0) put(1 + slot, doff, b * Rn + b, C_SW, ups_exp ? wb * VB[b][1] : wb);
      else put(1 + slot, doff, b * Rn + b, v == 1 ? C_SB : C_SD, wb * wb);
      for (int q = 0; q < Rn; ++q) {
        if (q == b) continue;
        if (v == 0) put(1 + slot, doff, b * Rn + q, C_G, VB[q][0] * prat(VB[b], VB[q]));
        else put(1 + slot, doff, b * Rn + q, v == 1 ? C_B1 : C_D1, wb * VB[q][0]);
      }
      for (int p = 0; p < Rn; ++p) {
        if (p == b) continue;
        if (v == 0) put(1 + slot, doff, p * Rn + b, C_G, VB[p][0] * prat(VB[p], VB[b]));
        else put(1 + slot, doff, p * Rn + b, v == 1 ? C_B2 : C_D2, VB[p][0] * wb);
      }
      ++slot;
    }
}
