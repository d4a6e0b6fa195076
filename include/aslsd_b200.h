/*
 * aslsd_b200 - C ABI of the B200 engine for the LSE hot path of saadlabyad/aslsd.
 *
 * Plain C: pointers and sizes only, no torch types.  Every entry point names
 * the reference interface (file:line under the reference tree) it replaces.
 * The caller owns every host buffer; the library owns device memory between
 * aslsd_create and aslsd_destroy.  All functions return 0 on success or a
 * negative code; the message is available from aslsd_last_error.  A context is
 * bound to one GPU (one process per GPU) and is not re-entrant.
 *
 * There is no CPU fallback: without a CUDA device aslsd_create fails.
 */
#ifndef ASLSD_B200_H
#define ASLSD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aslsd_ctx aslsd_ctx;

enum {
  ASLSD_OK = 0,
  ASLSD_ERR_CUDA = -1,      /* CUDA runtime error (no device, OOM, launch failure) */
  ASLSD_ERR_ARG = -2,       /* invalid argument / call order */
  ASLSD_ERR_UNSUPPORTED = -3 /* model outside the engine's scope (NotImplementedError upstream) */
};

/* Basis-kernel families.  The reference dispatches on str(type(.))
 * (functionals/kernels/basis_kernels/basis_kernel_exponential.py:418-420). */
enum {
  ASLSD_BASIS_EXPONENTIAL = 0, /* vars (omega, beta)        basis_kernel_exponential.py:387-439 */
  ASLSD_BASIS_GAUSSIAN = 1     /* vars (omega, beta, delta) basis_kernel_gaussian.py:448-529    */
};

/* One basis kernel of a KernelModel (functionals/kernels/kernel.py:14-92) with
 * its fixed variables (functionals/kernels/basis_kernel.py:113-130, :364-401). */
typedef struct {
  int32_t type;      /* ASLSD_BASIS_*                                             */
  int32_t free_mask; /* bit v set: variable v is a free parameter                 */
  double fixed[3];   /* value of variable v where bit v of free_mask is clear     */
} aslsd_basis;

/* Flags of the sum entry points (compute_f_sum / compute_diff_f_sum of
 * mhp_exact_estimator.py:709-711, grad_alloc of mhp_stratified_estimator.py:1333). */
enum {
  ASLSD_F = 1,          /* function sums                                          */
  ASLSD_DIFF = 2,       /* parameter-derivative sums                              */
  ASLSD_GRAD_ALLOC = 4  /* allocation statistic = ||grad||_2 instead of f          */
};

/* Options of aslsd_set_option. */
enum {
  ASLSD_OPT_LAG_QUADRATURE = 1, /* 1 (default): interior targets of every double sum go through the
                                  lag quadrature (one pass per pair of event streams);
                                  0: per-item pair kernels only */
  ASLSD_OPT_WEIGHT_CACHE = 2,  /* 1 (default): the quadrature weights - a function of the event streams
                                  and the lag windows only - are kept between calls while the events,
                                  the model, the target block and the non-omega parameters stay the
                                  same (the iterations of a fit); 0: recomputed on every call */
  ASLSD_OPT_ASYNC_EVENTS = 3,  /* page-locked inputs of aslsd_set_events are checked for order on the
                                  device, behind their transfer.  0 (default): aslsd_set_events waits
                                  for that verdict; 1: it returns as soon as the transfer is enqueued
                                  and a violation is reported (ASLSD_ERR_ARG, the context then has no
                                  events) by the next call that waits for the device - results of
                                  calls made in between are void */
  ASLSD_OPT_DEVICE_ASSEMBLY = 4 /* 1 (default): from the second call on the same row plan, the packed
                                  per-term sums are assembled from the raw rows by a device kernel and
                                  only they are read back; 0: raw rows read back, assembly on the host */
};

/* ---- lifetime ------------------------------------------------------------------ */
int aslsd_create(aslsd_ctx** out, int device);
int aslsd_destroy(aslsd_ctx* ctx);
const char* aslsd_last_error(const aslsd_ctx* ctx);
/* Use the caller's CUDA stream (a cudaStream_t) for all subsequent work;
 * NULL restores the context's own (non-blocking) stream.  To run on the legacy
 * default stream pass cudaStreamLegacy ((cudaStream_t)0x1). */
int aslsd_set_stream(aslsd_ctx* ctx, void* cuda_stream);
/* CUDA-event duration (ms) and launch count of the pair / sample kernels of
 * the last sum call, measured on the launching stream. */
int aslsd_last_timing(const aslsd_ctx* ctx, double* kernel_ms, int64_t* n_launches);
/* CUDA-event duration (ms) of the lag-moment kernel of the last exact-sum call (0 if it did
 * not run: quadrature off, no eligible stream, or cached weights reused). */
int aslsd_last_moment_timing(const aslsd_ctx* ctx, double* kernel_ms);

int aslsd_set_option(aslsd_ctx* ctx, int option, int value);

/* Measured fp64 FMA throughput of the device (TFLOP/s, 2 flops per DFMA): the
 * roofline denominator of the pair kernels (no reference counterpart). */
int aslsd_measure_fp64_peak(aslsd_ctx* ctx, double* tflops);

/* ---- events: replaces ProcessPath.__init__ data members ---------------------------
 * stats/events/process_path.py:63-106 (list_times, T_f, n_events, list_times2end).
 * times[i] is the sorted fp64 array of the n_events[i] events of type i.
 * Pageable arrays are copied into the context's own page-locked block (and checked for order)
 * by a few host threads and uploaded from there; the call returns with the transfers enqueued.
 * Arrays that are already page-locked (cudaHostAlloc / cudaHostRegister) are uploaded from
 * where they are, checked for order on the device (ASLSD_OPT_ASYNC_EVENTS) and read in place
 * by the host-side planning: they must stay valid and unchanged until the next
 * aslsd_set_events or aslsd_destroy. */
int aslsd_set_events(aslsd_ctx* ctx, int d, const int64_t* n_events,
                     const double* const* times, double T_f);
/* Move the end of the observation window of the events already on the device (ProcessPath.T_f,
 * list_times2end = T_f - t: process_path.py:70-76) without a new upload.  Sums evaluated afterwards see
 * psi(T_f - t_m), upsilon(T_f - t_m, .) at the new horizon; the caller restricts the targets to the
 * events before it (m_lo / m_hi).  Used for baselines that are constant on time intervals
 * (PieceConstBaseline.make_K, functionals/baselines/basis_baselines/basis_baseline_piececonst.py:172-254:
 * the K term is a combination of psi sums up to the interval ends). */
int aslsd_set_horizon(aslsd_ctx* ctx, double T_f);

/* ---- index maps: replaces time_ordering.get_kappa_varpi / get_lag_sizes ----------
 * stats/events/time_ordering.py:9-84 and :87-119.  Two-pass: query lengths,
 * then fill caller buffers.  All arrays are indexed [i*d + j].
 *   kappa[i][j] : n_events[j] + 1 entries (terminal entry per :64-65)
 *   varpi[i][j] : kappa[j][i][n_events[i]-1] + 2 entries (i != j), n_events[i] (i == j)
 *   lag_sizes[i][j] : same length as varpi[i][j]; [0] = 0, [h] = n_events[i] - varpi[i][j][h]
 * Any of the three output tables may be NULL to skip it. */
int aslsd_index_map_sizes(aslsd_ctx* ctx, int64_t* kappa_len, int64_t* varpi_len);
int aslsd_index_maps(aslsd_ctx* ctx, int64_t* const* kappa_out,
                     int64_t* const* varpi_out, int64_t* const* lag_sizes_out);

/* ---- model: replaces MHP.init_estimator / make_functionals ------------------------
 * models/hawkes/linear/mhp.py:1525-1538, :475-540.  n_basis[k*d + i] bases of
 * kernel phi_ki, concatenated in `bases` in row-major (k, i) order.  The
 * parameter vector x_k of sub-problem k is [mu_k, params(phi_k0), ...] with the
 * free variables of every basis in variable order (mhp.py:276-299, kernel.py:97-117). */
int aslsd_set_model(aslsd_ctx* ctx, int d, const int32_t* n_basis,
                    const aslsd_basis* bases);
/* Number of parameters of x_k and of each kernel phi_ki (n_param[i], i < d). */
int aslsd_model_n_param(const aslsd_ctx* ctx, int k, int32_t* n_param_k,
                        int32_t* n_param);

/* ---- exact sums: replaces MHPExactEstim.compute_objective -------------------------
 * optimize/estimators/mhp_exact_estimator.py:709-761 and the five compute_*_sum
 * methods :543-707.  Output `out` is packed term after term in the order the
 * reference evaluates them (psi_i, upsilonzero_i, phi_i, upsilon_i_i for
 * i = 0..d-1, then upsilon_i_j for i != j in itertools.product order); each term
 * holds [sum_f, sum_diff_f(ix_func 1)[n1], sum_diff_f(ix_func 2)[n2]] where
 * (n1, n2) = (n_param_ki, 1) for psi, (n_param_ki, 0) for upsilonzero / phi /
 * upsilon_i_i and (n_param_ki, n_param_kj) for upsilon_i_j.
 * m_lo / m_hi (length d, may be NULL for all events) restrict the *target*
 * events of every term to a block of indices per dimension: a time-block shard;
 * shard results add up to the full sums.
 * n_pair_terms (may be NULL, one entry per term): number of in-window event
 * pairs evaluated for the term (the unit of BASELINE.json's metric). */
int aslsd_exact_out_len(const aslsd_ctx* ctx, int k, int64_t* out_len, int64_t* n_terms);
int aslsd_exact_sums(aslsd_ctx* ctx, int k, const double* x_k, int flags,
                     const int64_t* m_lo, const int64_t* m_hi, double* out,
                     int64_t* n_pair_terms);

/* ---- per-event sums: replaces the O(N^2) loops of MHP.get_intensity_at_jumps ----------
 * (models/hawkes/linear/mhp.py:669-690) and of get_residuals_k_mhp (stats/residual_analysis/
 * goodness_of_fit.py:169-228).  out is d arrays of n_events[k] doubles, one after the other:
 *   kind 0: out[i][m] = sum over the type-i events before t^k_m of phi_ki(t^k_m - t^i_n)
 *   kind 1: out[i][m] = the same sum of psi_ki (kernel part of the compensator at t^k_m). */
int aslsd_point_sums(aslsd_ctx* ctx, int k, const double* x_k, int kind, double* out);

/* ---- host-side index sampling: replaces utilities/useful_statistics.py:38-127 ---------
 * sample_single_ix_stratified / sample_double_ix_stratum / sample_double_ix_stratified draw with
 * numpy's Generator(PCG64); these entry points restate that generator and Generator.choice bit
 * for bit, start from the caller's generator state and return the advanced state, so one native
 * call replaces the thousands of rng.choice calls of a term without changing a single draw.
 * State layout: numpy's rng.bit_generator.state ('state', 'inc' as two 64-bit halves each,
 * 'has_uint32', 'uinteger').  Host code only; no device is touched. */
typedef struct {
  uint64_t state_hi, state_lo, inc_hi, inc_lo;
  int32_t has_uint32;
  uint32_t uinteger;
} aslsd_pcg64;
/* Generator.choice(pop, size, replace=replace) (no p, shuffle=True) into out[size]. */
int aslsd_rng_choice(aslsd_pcg64* gen, int64_t pop, int64_t size, int replace, int64_t* out);
/* All strata of one single-sum term (utilities/useful_statistics.py:38-60,
 * sample_single_ix_stratified): stratum s covers the indices [lo[s], hi[s]] and gets
 * n_samples[s] of them without replacement, written one stratum after the other to out. */
int aslsd_sample_single_ix(aslsd_pcg64* gen, int64_t n_strata, const int64_t* lo, const int64_t* hi,
                           const int64_t* n_samples, int64_t* out);
/* All strata of one double-sum term (targets of type i, sources of type j): stratum s covers
 * the index lags [h_inf[s], h_sup[s]] and gets n_samples[s] pairs, written one stratum after
 * the other to m_out / n_out.  varpi_ij = varpi[i][j], kappa_ji = kappa[j][i]. */
int aslsd_sample_double_ix(aslsd_pcg64* gen, int64_t n_strata, const int64_t* h_inf,
                           const int64_t* h_sup, const int64_t* n_samples, const int64_t* varpi_ij,
                           const int64_t* kappa_ji, int64_t n_events_i, int64_t* m_out, int64_t* n_out);

/* ---- marked models: replaces the pair sums of MTLHExactEstim.compute_objective -------
 * optimize/estimators/mtlh_exact_estimator.py:685-971 for constant baselines.  The impact
 * functions stay the reference's Python (functionals/impact_functions/impact_function.py:
 * 116-131): the caller evaluates I_ki(xi^i) and its parameter derivatives per event and hands
 * them over as weight vectors; vector v holds one double per event of type dims[v].
 * aslsd_exact_sums_weighted computes the terms of aslsd_exact_sums with every event pair /
 * event weighted by the vectors that `roles` names (5 * d entries, -1 = unweighted):
 *   roles[0*d+i] sources of phi_i, [1*d+i] targets of upsilon_i_*, [2*d+i] sources of
 *   upsilon_*_i, [3*d+i] targets of psi_i (the K_i term), [4*d+i] targets of upsilonzero_i.
 * term_mask (may be NULL): one byte per term in output order, 0 = term left at zero. */
int aslsd_set_weights(aslsd_ctx* ctx, int n_vec, const int32_t* dims, const double* const* vals);
int aslsd_exact_sums_weighted(aslsd_ctx* ctx, int k, const double* x_k, int flags,
                              const int64_t* m_lo, const int64_t* m_hi, const int32_t* roles,
                              const uint8_t* term_mask, double* out, int64_t* n_pair_terms);

/* Several sub-problems in one call: what d calls of lse_k_grad_estimate compute
 * (the k loop of MHP.fit, models/hawkes/linear/mhp.py:1771-1787, or MHP.get_lse,
 * :639-664).  Work items whose raw sums coincide - same event streams, target
 * block, basis (beta, delta) and component mask; omega never enters - are
 * evaluated once for the whole batch.  outs[q] / n_pair_terms[q] as in
 * aslsd_exact_sums for sub-problem ks[q] with parameters xs[q]. */
int aslsd_exact_sums_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs,
                           int flags, const int64_t* m_lo, const int64_t* m_hi,
                           double* const* outs, int64_t* const* n_pair_terms);

/* Split form for multi-GPU: device-side raw accumulators (to be summed across
 * ranks with one allreduce) and host-side assembly of the packed terms.
 * raw_len is an upper bound that does not depend on the parameters. */
int aslsd_exact_raw_len(aslsd_ctx* ctx, int k, int64_t* raw_len);
int aslsd_exact_raw_device(aslsd_ctx* ctx, int k, const double* x_k, int flags,
                           const int64_t* m_lo, const int64_t* m_hi,
                           double* raw_device, int64_t* n_pair_terms);
int aslsd_exact_assemble(aslsd_ctx* ctx, int k, const double* x_k, int flags,
                         const double* raw_host, double* out);
/* Largest lag any double sum of the batch reaches at these parameters (needs the model, not the
 * events): the halo of earlier events a time-block shard must hold before its first target. */
int aslsd_exact_max_lag(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs, int flags,
                        double* max_lag);
/* Doubles of the raw buffer written by the last raw_device call (a prefix of raw_len): the
 * part to all-reduce and to copy to the host. */
int aslsd_exact_raw_used(aslsd_ctx* ctx, int64_t* raw_used);
int aslsd_exact_raw_len_batch(aslsd_ctx* ctx, int n_k, int64_t* raw_len);
int aslsd_exact_raw_device_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs,
                                 int flags, const int64_t* m_lo, const int64_t* m_hi,
                                 double* raw_device);
int aslsd_exact_assemble_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs,
                               int flags, const double* raw_host, double* const* outs);
/* The same with the assembly on the device: the packed sums of the batch (the sub-problems
 * concatenated in call order, aslsd_exact_packed_len_batch doubles in all; layout of each as in
 * aslsd_exact_sums) are left in packed_device (device memory of the caller, e.g. a torch tensor) on
 * the context's stream.  recombine_* is linear (mhp_exact_estimator.py:442-530), so the shards of a
 * time-block partition all-reduce these few thousand doubles instead of the raw rows. */
int aslsd_exact_packed_len_batch(aslsd_ctx* ctx, int n_k, const int* ks, int64_t* packed_len);
int aslsd_exact_packed_device_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs,
                                    int flags, const int64_t* m_lo, const int64_t* m_hi,
                                    double* packed_device);
/* Diagnostic: the linear map of the device assembly evaluated on the host, in the kernel's order of
 * operations - same arguments and results as aslsd_exact_assemble_batch (no device work). */
int aslsd_exact_assemble_map_batch(aslsd_ctx* ctx, int n_k, const int* ks, const double* const* xs,
                                   int flags, const double* raw_host, double* const* outs);
/* ---- simulation ---------------------------------------------------------------------------
 * Cluster (branching) simulation of a linear Hawkes process on the device: the generation loop of
 * MHP.simulate (models/hawkes/linear/mhp.py:733-854).  Same law as the reference, not its random
 * numbers (counter-based Philox streams keyed by `seed`).  bases: the d x d kernels phi_ij in
 * row-major order (child type i, parent type j), n_basis[i * d + j] entries each, every basis with
 * all its variables in `fixed` (omega, beta[, delta]; free_mask ignored); mu: d baseline rates.
 * aslsd_simulate leaves the sorted event times on the device and returns the counts per type;
 * aslsd_simulate_fetch copies them into the caller's arrays (n_events[i] doubles) and frees the pool. */
int aslsd_simulate(aslsd_ctx* ctx, int d, const int32_t* n_basis, const aslsd_basis* bases,
                   const double* mu, double T_f, uint64_t seed, int64_t* n_events);
int aslsd_simulate_fetch(aslsd_ctx* ctx, double* const* times);

/* Wait for everything enqueued on the context's stream. */
int aslsd_synchronize(aslsd_ctx* ctx);

/* Closed-form evaluations performed by the last exact-sum call (basis-pair
 * evaluations of the pair kernels / of the single-index kernels): the work
 * count behind the roofline figures. */
int aslsd_last_eval_counts(aslsd_ctx* ctx, int64_t* pair_evals, int64_t* single_evals);
/* Full work accounting of the last exact-sum call, counts[0..7] (n >= 8):
 *   [0] basis-pair evaluations of the per-item pair kernels   [1] single-index evaluations
 *   [2] event pairs visited by the shared-class kernel
 *   [3] exp, [4] normcdf, [5] multiply-add evaluations of the shared-class kernel
 *   [6] unique work items evaluated   [7] work items before de-duplication */
int aslsd_last_work_counts(aslsd_ctx* ctx, int64_t* counts, int n);

/* What every row of the raw accumulator array of the last batch plan sums (after a
 * raw_device / exact_sums call): row r adds up, over its event pairs (pair families) or
 * target events (single-index families), the components of one basis pair / one basis with the
 * variables params[11 r ..] = (beta_1, delta_1, beta_2, delta_2, beta_own, delta_own, the
 * truncation of the row: lags outside [s_lo, s_hi] are not visited, targets whose time to T_f is
 * at least t_int take the closed forms' limits at infinity, and the block [m_lo, m_hi) of target
 * events the call was asked for);
 * meta[7 r ..] = (family, target event type, source event type, same-type flag, component
 * mask, weight vector of the targets, weight vector of the sources: indices into
 * aslsd_set_weights, -1 = unweighted).  Family codes: 0 / 1 upsilon exp x exp / gauss x gauss, 2 / 3 phi exp / gauss,
 * 4 / 5 psi exp / gauss, 6 / 7 upsilon at lag zero exp / gauss, 8-11 the mixed pairs.
 * n_rows receives the number of rows; at most cap rows are written.  Host-side diagnostic:
 * with it the planning (which parameters and event streams enter which row) and the assembly of
 * the packed sums (aslsd_exact_assemble_batch) are checked on a GPU-less machine against the
 * oracle, tests/test_plan_and_assembly_host.py. */
int aslsd_exact_plan_rows(aslsd_ctx* ctx, int64_t cap, int64_t* n_rows, int32_t* meta, double* params);

/* ---- sampled sums: replaces the bodies of MHPStratEstim.estimate_sum_* ------------
 * optimize/estimators/mhp_stratified_estimator.py:534-900, for every term of one
 * compute_objective call (:1333-1386) in a single launch.
 * The (m, n) samples are generated on the host by the reference's own RNG code
 * (utilities/useful_statistics.py:38-127) and consumed here unchanged; the index
 * -> time transforms are those of the compute_*_sum closures (:1097-1100,
 * :1148-1151, :1196-1203, :1245-1252, :1300-1307).
 *   terms[q]: a named term and the number of its strata in this batch; strata
 *             of all terms are concatenated in order.
 *     kind: 0 psi_i, 1 upsilonzero_i, 2 phi_i, 3 upsilon_i_j (i == j allowed)
 *   stratum_offsets: n_strata_total + 1 offsets into ix_m / ix_n
 *   ix_n is ignored for single-index terms (psi, upsilonzero)
 * Per stratum s of a term with n_out = 1 + n1 + n2 outputs (as in aslsd_exact_sums;
 * aslsd_sampled_n_out):
 *   sum_out[off_s + c]      plain sum over the stratum's samples of f (c = 0) and
 *                           of every derivative; off_s = running sum of n_out
 *                           (the mean numpy.mean is sum / count)
 *   stat_mean[s], stat_std[s]  mean and population std (numpy.std, ddof = 0) of
 *                           the allocation statistic: f, or ||grad f||_2 when
 *                           ASLSD_GRAD_ALLOC is set (:605-635)
 * Fused allocation update (:637-645), when strata_sizes, default_alloc and
 * rel_alloc are given: maximal runs of consecutive strata of a term with
 * strata_sizes[s] > 0 form one adaptive group;
 *   rel_alloc[s] = size_s * sqrt(size_s / (size_s - 1)) * std_s, normalised over
 * the group, or default_alloc[s] when every std of the group is zero.  Strata
 * outside any group get default_alloc[s]. */
/* Saturated samples of the single-index terms (psi_i, upsilonzero_i): for a target further than
 * sat_t from T_f every upper integration limit has reached +infinity in fp64, and the reference's
 * per-sample value is one constant - computed by the caller with the reference's own closure
 * (model.psi[k][i] / model.upsilon[i][i][k] at a large t) and handed in as v_sat, so that the
 * samples of such strata carry the reference's value bit for bit (its "standard deviation" there
 * is the rounding residue of numpy's summation, and the next integer allocations depend on it,
 * :631-645).  sat_t = +infinity (or NaN) switches this off; v_sat only replaces the function value. */
typedef struct {
  int32_t kind;
  int32_t i, j;
  int32_t n_strata;
  double sat_t;
  double v_sat;
} aslsd_sample_term;

int aslsd_sampled_n_out(const aslsd_ctx* ctx, int k, int term_kind, int i, int j, int32_t* n_out);
int aslsd_sampled_sums(aslsd_ctx* ctx, int k, const double* x_k, int flags, int n_terms,
                       const aslsd_sample_term* terms, const int64_t* stratum_offsets,
                       const int64_t* ix_m, const int64_t* ix_n, const double* strata_sizes,
                       const double* default_alloc, double* sum_out, double* stat_mean,
                       double* stat_std, double* rel_alloc);

/* Sampled sums with the same weights (MTLHStratEstim.estimate_sum_*dom,
 * optimize/estimators/mtlh_stratified_estimator.py:615-1025): arguments of aslsd_sampled_sums
 * plus `roles`; every per-sample output is multiplied by weight(target) * weight(source). */
int aslsd_sampled_sums_weighted(aslsd_ctx* ctx, int k, const double* x_k, int flags, int n_terms,
                                const aslsd_sample_term* terms, const int64_t* stratum_offsets,
                                const int64_t* ix_m, const int64_t* ix_n, const double* strata_sizes,
                                const double* default_alloc, const int32_t* roles, double* sum_out,
                                double* stat_mean, double* stat_std, double* rel_alloc);

/* Stand-alone allocation update of estimate_sum_adaptivedom (:637-645) for one
 * group of strata. */
int aslsd_update_allocation(aslsd_ctx* ctx, int n_strata, const double* strata_sizes,
                            const double* stat_std, const double* default_alloc,
                            double* rel_alloc_out);

#ifdef __cplusplus
}
#endif
#endif /* ASLSD_B200_H */
