// Shared host-side state of the aslsd_b200 engine (one context per GPU).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <array>
#include <string>
#include <vector>

#include "../../include/aslsd_b200.h"

#define ASLSD_MAX_COMP 8       // raw accumulators per work item
#define ASLSD_BUCKET_SHIFT 2   // one time bucket per 2^2 events (lookup table)

struct BasisHost {
  int type;
  int free_mask;
  double fixed[3];
  int n_vars() const { return type == ASLSD_BASIS_EXPONENTIAL ? 2 : 3; }
  int n_param() const {
    int n = 0;
    for (int v = 0; v < n_vars(); ++v) n += (free_mask >> v) & 1;
    return n;
  }
};

struct KernelHost {
  std::vector<BasisHost> bases;
  int n_param() const {
    int n = 0;
    for (auto& b : bases) n += b.n_param();
    return n;
  }
};

// Device-side description of one work item of the pair / single kernels.
// An item is one basis pair (upsilon), one basis (phi, psi) of one named term.
struct ItemDev {
  int32_t family;    // ITEM_* below
  int32_t mask;      // which raw components to accumulate
  int32_t tgt_dim, src_dim;
  int32_t same;      // target and source are the same event type (n < m by index)
  int32_t n_chunks;  // work units of this item
  int64_t m_lo, m_hi;    // target block
  int64_t tpw;           // targets per work unit
  int64_t wu_first;      // first work unit (prefix over items)
  double s_lo, s_hi;     // lag window
  double t_int;          // targets with T - t_m >= t_int are interior
  double b1, d1, b2, d2; // beta / delta of the two bases (cross components)
  double bo, dlo;        // beta / delta of the own basis (sim-own components)
  int64_t dst_row;       // row of the raw accumulator array this item adds into
  int64_t n_int;         // single-index rows: interior targets taken as n_int x closed form
  int64_t wa_off, wb_off; // per-event weights of the targets / sources: offsets (in doubles) into the
                          // context's weight block, -1 = no weight (marked models, aslsd_set_weights)
};

enum {
  ITEM_UPS_EXP = 0,
  ITEM_UPS_GAUSS = 1,
  ITEM_PHI_EXP = 2,
  ITEM_PHI_GAUSS = 3,
  ITEM_PSI_EXP = 4,
  ITEM_PSI_GAUSS = 5,
  ITEM_UPS0_EXP = 6,   // upsilon at lag zero: one evaluation per target
  ITEM_UPS0_GAUSS = 7,
  // mixed pairs: basis 1 exponential x basis 2 Gaussian (EG) and the converse (GE)
  ITEM_UPS_EG = 8,
  ITEM_UPS_GE = 9,
  ITEM_UPS0_EG = 10,
  ITEM_UPS0_GE = 11,
  ITEM_N_FAMILIES = 12
};

__host__ __device__ inline bool item_pair_family(int f) {
  return f == ITEM_UPS_EXP || f == ITEM_UPS_GAUSS || f == ITEM_PHI_EXP || f == ITEM_PHI_GAUSS ||
         f == ITEM_UPS_EG || f == ITEM_UPS_GE;
}

// Raw component slots.
//   upsilon exp  : 0 G, 1 B1, 2 B2, 5 SW, 6 SB
//   upsilon gauss: 0 G, 1 B1, 2 D1, 3 B2, 4 D2, 5 SW, 6 SB, 7 SD
//   upsilon exp x gauss: 0 G, 1 B1, 3 B2, 4 D2, 5 G with the reference's overflow guard
//   upsilon gauss x exp: 0 G, 1 B1, 2 D1, 3 B2, 5 G with the reference's overflow guard
//   phi exp      : 0 E0, 1 EB            phi gauss: 0 P, 1 PB, 2 PD
//   psi exp      : 0 I, 1 IB             psi gauss: 0 I, 1 IB, 2 ID
enum { C_G = 0, C_B1 = 1, C_D1 = 2, C_B2 = 3, C_D2 = 4, C_SW = 5, C_SB = 6, C_SD = 7 };
// exponential upsilon reuses slots: B1 -> 1, B2 -> 3
#define MASK(c) (1 << (c))
#define MASK_OWN_SAME (1 << 8)   // own basis has the same (beta, delta) as both cross bases

struct TermPlan {
  int kind;          // 0 psi, 1 upsilonzero, 2 phi, 3 upsilon
  int i, j;          // dimensions (j == i for single-index terms and upsilon_i_i)
  int n1, n2;        // derivative counts (packed output)
  int64_t out_off;   // offset in the packed output
  int item_first, item_count;   // items computing this term's cross components
};

struct aslsd_ctx;
// The context's compute stream.  Event uploads from page-locked arrays run on a copy stream of their
// own (ctx.cu: aslsd_set_events); whoever takes the compute stream as a cudaStream_t - every kernel
// launch, copy and event record written `ctx->stream` - first orders it behind those uploads
// (aslsd_join_uploads).  Only the launch sequence that knows what it may start before the last event
// type has arrived (exact.cu: the moment kernels of the streams whose types are there) asks for raw().
struct CtxStream {
  cudaStream_t s = nullptr;
  aslsd_ctx* owner = nullptr;
  inline operator cudaStream_t() const;
  cudaStream_t raw() const { return s; }
};

struct aslsd_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  CtxStream stream;
  // uploads in flight on the copy stream: ev_type[i] = type i (and every type before it) is on the
  // device, checked for order, its bucket table built
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> ev_type;
  int n_upload_types = 0;
  bool uploads_pending = false;
  size_t upload_bytes = 0;              // of the uploads in flight
  // second lane of the moment kernels launched behind those uploads (exact.cu: launch_moments)
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_aux0 = nullptr, ev_aux1 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // second lane of an exact-sum call (exact.cu: run_batch_impl)
  std::string err;

  // events
  int d = 0;
  double T_f = 0.;
  std::vector<int64_t> n_events;
  std::vector<int64_t> off;        // offsets (in doubles) into d_times, 16 B aligned
  double* d_times = nullptr;       // all event types, padded
  int64_t times_len = 0;
  // per-dimension time-bucket lookup tables (lower_bound accelerators)
  std::vector<int64_t> tab_off;    // offsets into d_tab
  std::vector<int64_t> tab_len;    // number of buckets per dim
  std::vector<double> tab_t0, tab_w;
  int32_t* d_tab = nullptr;
  // device copies of the per-dimension metadata
  void* d_dims = nullptr;          // DimDev[d]

  // per-event weight vectors (impact values of marked models) and how the next plan uses them
  double* d_weights = nullptr; size_t weights_cap = 0;
  std::vector<int64_t> weight_off;    // offset of vector v in d_weights
  std::vector<int> weight_dim;        // event type of vector v
  const int32_t* roles = nullptr;     // [5 * d] weight vector per (role, type), see aslsd_exact_sums_weighted
  const uint8_t* term_mask = nullptr; // per term: 0 = skip

  // model
  bool has_model = false;
  std::vector<KernelHost> kernels;   // [k*d + i]

  // scratch
  void* d_items = nullptr;  size_t items_cap = 0;
  double* d_partials = nullptr; size_t partials_cap = 0;
  double* d_raw = nullptr; size_t raw_cap = 0;
  unsigned long long* d_counts = nullptr; size_t counts_cap = 0;
  unsigned long long* d_qcounts = nullptr; size_t qcounts_cap = 0;   // pair counts of the moment kernels' work units
  void* h_pinned = nullptr; size_t pinned_cap = 0;
  // page-locked staging of the plan tables (items, streams, rows, boundary groups); ev_tables marks
  // the end of their last transfer, so the block is rewritten only once the device has read it
  void* h_tables = nullptr; size_t tables_cap = 0;
  cudaEvent_t ev_tables = nullptr; bool tables_in_flight = false;
  void* d_sample = nullptr; size_t sample_cap = 0;

  // host copy of the event times (planning), inside one pinned staging block that is
  // also the source of the host -> device copy
  double* h_events = nullptr; size_t h_events_cap = 0;
  double* h_pad = nullptr;         // a few page-locked pad values (uploads from the caller's own pinned arrays)
  // order check of page-locked inputs: done on the device behind the transfer; the verdict
  // (h_flag, valid once ev_flag has completed) is collected by events_verdict()
  int* d_flag = nullptr;
  int* h_flag = nullptr;
  cudaEvent_t ev_flag = nullptr;
  bool verdict_pending = false;
  bool async_events = false;       // ASLSD_OPT_ASYNC_EVENTS
  std::vector<const double*> h_times;
  std::vector<unsigned char> h_dims;   // host copy of the DimDev table (source of an async upload)
  size_t times_cap = 0, tab_cap = 0, dims_cap = 0;   // capacities of d_times / d_tab / d_dims
  // lag quadrature (lag_quadrature.cuh): stream / row tables, partial moments, node weights
  void* d_qtab = nullptr; size_t qtab_cap = 0;
  double* d_q_partials = nullptr; size_t q_partials_cap = 0;
  double* d_q_weights = nullptr; size_t q_weights_cap = 0;
  bool use_quadrature = true;
  bool cache_weights = true;        // reuse the node weights while plan and events stay the same
  bool last_weights_reused = false;
  // geometry of the streams whose node weights are in d_q_weights (they depend on the events and
  // the bin grids only: plans of other sub-problems with the same streams reuse them)
  std::vector<unsigned char> w_streams;    // QStreamDev records
  int64_t w_events_epoch = -1;
  bool lq_tables_set = false;
  size_t q_smem_set = 0;
  // last batch plan (exact.cu): reused while the model, the events, the flags, the
  // target block and the non-omega parameters stay the same (plans hold no results)
  void* plan_cache = nullptr;
  void (*plan_cache_free)(void*) = nullptr;
  // index maps (index_maps.cu): pair table, lag counts and the three tables, kept between calls
  void* d_im_pairs = nullptr; size_t im_pairs_cap = 0;
  int64_t* d_im_H = nullptr; size_t im_H_cap = 0;
  int64_t* d_im_kappa = nullptr; size_t im_kappa_cap = 0;
  int64_t* d_im_varpi = nullptr; size_t im_varpi_cap = 0;
  int64_t* d_im_lags = nullptr; size_t im_lags_cap = 0;
  // device assembly of the packed sums (exact.cu; the CSR maps belong to the plans): omega tables, result
  double* d_wtab = nullptr; size_t wtab_cap = 0;
  double* h_wtab = nullptr; size_t h_wtab_cap = 0;     // page-locked, two halves used in turn
  int wtab_flip = 0;
  bool wtab_in_flight[2] = {false, false};
  cudaEvent_t ev_wtab[2] = {nullptr, nullptr};
  const void* tables_owner = nullptr;                  // the plan whose item / stream tables are on the device
  int64_t tables_owner_uid = 0, plan_uid = 0;
  double* d_packed = nullptr; size_t packed_cap = 0;
  bool device_assembly = true;                         // ASLSD_OPT_DEVICE_ASSEMBLY
  // device-side simulation (simulate.cu): the event pool of the last aslsd_simulate until it is fetched
  void* sim_state = nullptr;
  void (*sim_state_free)(void*) = nullptr;
  int64_t model_epoch = 0;    // bumped by set_model / set_option
  int64_t events_epoch = 0;   // bumped by set_events
  int64_t last_work[8] = {0};   // see aslsd_last_work_counts
  int64_t last_n_wu_q = 0, last_n_bgroups = 0;
  // timing
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;     // around all kernels of a sum call
  cudaEvent_t ev_q0 = nullptr, ev_q1 = nullptr; // around the lag-moment kernel
  double last_ms = 0., last_moment_ms = 0.;
  bool timing_pending = false, moment_timing_pending = false;
  int64_t last_launches = 0;
  // evaluation counters of the last exact-sum call (per item, on the device)
  int64_t last_n_wu = 0;
  int last_n_items = 0;
  int last_fam_first[ITEM_N_FAMILIES] = {0}, last_fam_count[ITEM_N_FAMILIES] = {0};
};

inline void aslsd_join_uploads(aslsd_ctx* ctx) {
  if (!ctx->uploads_pending) return;
  ctx->uploads_pending = false;
  if (ctx->n_upload_types > 0) cudaStreamWaitEvent(ctx->stream.s, ctx->ev_type[(size_t)ctx->n_upload_types - 1], 0);
}
inline CtxStream::operator cudaStream_t() const {
  if (owner) aslsd_join_uploads(owner);
  return s;
}


struct DimDev {
  const double* t;     // event times of this type
  const int32_t* tab;  // bucket table: tab[g] = lower_bound(t, t0 + g * w), g = 0..n_buckets
  int64_t n;
  int64_t n_buckets;
  double t0, w, inv_w;
  const int* bad;      // verdict of the device-side order check of this upload (shared by all types):
                       // kernels whose indexing relies on sorted times return at once when it is set
};

int aslsd_fail(aslsd_ctx* ctx, int code, const std::string& msg);
// Collect the verdict of a device-side order check (see aslsd_set_events).  wait = false: only if
// it is already there.  On a violation the context drops its events and the error is returned.
int events_verdict(aslsd_ctx* ctx, bool wait);
#define CUDA_TRY(ctx, expr)                                                        \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess)                                                         \
      return aslsd_fail(ctx, ASLSD_ERR_CUDA,                                       \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));       \
  } while (0)

int ensure_device_buffer(aslsd_ctx* ctx, void** ptr, size_t* cap, size_t bytes);
int ensure_pinned_buffer(aslsd_ctx* ctx, size_t bytes);

// lower_bound through the bucket table: first index with t[idx] >= x
__device__ __forceinline__ int64_t dim_lower_bound(const DimDev& D, double x) {
  if (D.n == 0) return 0;
  double gf = (x - D.t0) * D.inv_w;
  int64_t g = gf <= 0. ? 0 : (gf >= (double)(D.n_buckets - 1) ? D.n_buckets - 1 : (int64_t)gf);
  // make the bucket choice consistent with the boundaries used at build time
  while (g > 0 && x < D.t0 + (double)g * D.w) --g;
  while (g < D.n_buckets - 1 && x >= D.t0 + (double)(g + 1) * D.w) ++g;
  int64_t lo = D.tab[g], hi = D.tab[g + 1];
  if (g == 0 && x < D.t0) { lo = 0; }
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (D.t[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// first index with t[idx] > x
__device__ __forceinline__ int64_t dim_upper_bound(const DimDev& D, double x) {
  int64_t p = dim_lower_bound(D, x);
  while (p < D.n && D.t[p] == x) ++p;
  return p;
}
