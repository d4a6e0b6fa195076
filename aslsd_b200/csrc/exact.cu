// Exact LSE sums of one sub-problem k: launches and assembly.  Replaces
// MHPExactEstim.compute_objective and its five compute_*_sum methods
// (optimize/estimators/mhp_exact_estimator.py:543-761).
#include "pair_kernels.cuh"
#include "plan.h"

#include <cmath>
#include <cstring>

template <int FAMILY>
static void launch_pair(aslsd_ctx* ctx, const SumPlan& P, const ItemDev* d_items) {
  if (P.fam_wu_count[FAMILY] == 0) return;
  pair_kernel<FAMILY><<<(unsigned)P.fam_wu_count[FAMILY], PAIR_NT, 0, ctx->stream>>>(
      (const DimDev*)ctx->d_dims, d_items + P.fam_first[FAMILY], P.fam_count[FAMILY], ctx->T_f,
      ctx->d_partials, ctx->d_counts, P.fam_wu_first[FAMILY]);
  ctx->last_launches++;
}
template <int FAMILY>
static void launch_single(aslsd_ctx* ctx, const SumPlan& P, const ItemDev* d_items) {
  if (P.fam_wu_count[FAMILY] == 0) return;
  single_kernel<FAMILY><<<(unsigned)P.fam_wu_count[FAMILY], PAIR_NT, 0, ctx->stream>>>(
      (const DimDev*)ctx->d_dims, d_items + P.fam_first[FAMILY], P.fam_count[FAMILY], ctx->T_f,
      ctx->d_partials, ctx->d_counts, P.fam_wu_first[FAMILY]);
  ctx->last_launches++;
}

// Launch everything; leaves raw[item][comp] in raw_device.
static int run_plan(aslsd_ctx* ctx, const SumPlan& P, double* raw_device, unsigned long long* d_item_counts) {
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

// ---- assembly: raw components -> packed per-term sums ---------------------------------------
static void assemble(const aslsd_ctx* ctx, const SumPlan& P, const double* x_k, const double* raw,
                     double* out) {
  for (int64_t c = 0; c < P.out_len; ++c) out[c] = 0.;
  for (size_t t = 0; t < P.terms.size(); ++t) {
    double* o = out + P.terms[t].out_off;
    emit_term(ctx, P, t, x_k, [&](int c, int item, int comp, double coef, double outer) {
      o[c] += outer * (coef * raw[(size_t)item * ASLSD_MAX_COMP + comp]);
    });
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
  SumPlan P;
  int rc = build_plan(ctx, k, x_k, flags, m_lo, m_hi, false, P);
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
  SumPlan P;
  // the item table (hence the raw layout) depends only on the model, the flags
  // and the parameters, not on the target block
  int rc = build_plan(ctx, k, x_k, flags, nullptr, nullptr, false, P);
  if (rc) return rc;
  assemble(ctx, P, x_k, raw_host, out);
  return ASLSD_OK;
}

extern "C" int aslsd_exact_sums(aslsd_ctx* ctx, int k, const double* x_k, int flags, const int64_t* m_lo,
                                const int64_t* m_hi, double* out, int64_t* n_pair_terms) {
  if (!ctx || !x_k || !out) return aslsd_fail(ctx, ASLSD_ERR_ARG, "exact_sums: bad arguments");
  SumPlan P;
  int rc = build_plan(ctx, k, x_k, flags, m_lo, m_hi, false, P);
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
