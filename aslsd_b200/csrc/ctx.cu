// Context lifetime, event upload and model description.
//   aslsd_set_events  <- stats/events/process_path.py:63-106 (data members)
//   aslsd_set_model   <- models/hawkes/linear/mhp.py:1525-1538 (init_estimator)
#include "common.cuh"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>

#ifndef ASLSD_STAGE_THREADS
#define ASLSD_STAGE_THREADS 8
#endif

int aslsd_fail(aslsd_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

int ensure_device_buffer(aslsd_ctx* ctx, void** ptr, size_t* cap, size_t bytes) {
  if (bytes <= *cap && *ptr) return ASLSD_OK;
  if (*ptr) CUDA_TRY(ctx, cudaFree(*ptr));
  *ptr = nullptr;
  *cap = 0;
  size_t want = std::max<size_t>(bytes + bytes / 4, 256);
  CUDA_TRY(ctx, cudaMalloc(ptr, want));
  *cap = want;
  return ASLSD_OK;
}

int ensure_pinned_buffer(aslsd_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_cap && ctx->h_pinned) return ASLSD_OK;
  if (ctx->h_pinned) CUDA_TRY(ctx, cudaFreeHost(ctx->h_pinned));
  ctx->h_pinned = nullptr;
  ctx->pinned_cap = 0;
  size_t want = std::max<size_t>(bytes + bytes / 4, 4096);
  CUDA_TRY(ctx, cudaMallocHost(&ctx->h_pinned, want));
  ctx->pinned_cap = want;
  return ASLSD_OK;
}

extern "C" int aslsd_create(aslsd_ctx** out, int device) {
  if (!out) return ASLSD_ERR_ARG;
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0 || device < 0 || device >= n_dev) return ASLSD_ERR_CUDA;
  aslsd_ctx* ctx = new aslsd_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_aux0, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_aux1, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreate(&ctx->ev_q0) != cudaSuccess || cudaEventCreate(&ctx->ev_q1) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_flag, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_tables, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_wtab[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_wtab[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaMalloc((void**)&ctx->d_flag, sizeof(int)) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_flag, sizeof(int)) != cudaSuccess ||
      cudaMemset(ctx->d_flag, 0, sizeof(int)) != cudaSuccess) {
    delete ctx;
    return ASLSD_ERR_CUDA;
  }
  ctx->stream.s = ctx->own_stream;
  ctx->stream.owner = ctx;
  *out = ctx;
  return ASLSD_OK;
}

extern "C" int aslsd_destroy(aslsd_ctx* ctx) {
  if (!ctx) return ASLSD_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  for (auto ev : ctx->ev_type) cudaEventDestroy(ev);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
  if (ctx->ev_aux0) cudaEventDestroy(ctx->ev_aux0);
  if (ctx->ev_aux1) cudaEventDestroy(ctx->ev_aux1);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  cudaFree(ctx->d_times);
  cudaFree(ctx->d_tab);
  cudaFree(ctx->d_dims);
  cudaFree(ctx->d_items);
  cudaFree(ctx->d_partials);
  cudaFree(ctx->d_raw);
  cudaFree(ctx->d_counts);
  cudaFree(ctx->d_qcounts);
  if (ctx->plan_cache && ctx->plan_cache_free) ctx->plan_cache_free(ctx->plan_cache);
  if (ctx->sim_state && ctx->sim_state_free) ctx->sim_state_free(ctx->sim_state);
  cudaFree(ctx->d_sample);
  cudaFree(ctx->d_weights);
  cudaFree(ctx->d_qtab);
  cudaFree(ctx->d_q_partials);
  cudaFree(ctx->d_q_weights);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_events) cudaFreeHost(ctx->h_events);
  if (ctx->h_pad) cudaFreeHost(ctx->h_pad);
  if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
  if (ctx->h_tables) cudaFreeHost(ctx->h_tables);
  cudaEventDestroy(ctx->ev_tables);
  cudaFree(ctx->d_im_pairs);
  cudaFree(ctx->d_im_H);
  cudaFree(ctx->d_im_kappa);
  cudaFree(ctx->d_im_varpi);
  cudaFree(ctx->d_im_lags);
  cudaFree(ctx->d_wtab);
  cudaFree(ctx->d_packed);
  if (ctx->h_wtab) cudaFreeHost(ctx->h_wtab);
  cudaEventDestroy(ctx->ev_wtab[0]);
  cudaEventDestroy(ctx->ev_wtab[1]);
  cudaFree(ctx->d_flag);
  cudaEventDestroy(ctx->ev_flag);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->ev_q0);
  cudaEventDestroy(ctx->ev_q1);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return ASLSD_OK;
}

extern "C" const char* aslsd_last_error(const aslsd_ctx* ctx) {
  return ctx ? ctx->err.c_str() : "no context (no CUDA device?)";
}

extern "C" int aslsd_set_stream(aslsd_ctx* ctx, void* cuda_stream) {
  if (!ctx) return ASLSD_ERR_ARG;
  cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  if (next != ctx->stream.raw()) {
    // uploads and kernels already enqueued are ordered on the old stream only
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->stream.s = next;
  return ASLSD_OK;
}

extern "C" int aslsd_synchronize(aslsd_ctx* ctx) {
  if (!ctx) return ASLSD_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ASLSD_OK;
}

extern "C" int aslsd_last_timing(const aslsd_ctx* ctx_, double* kernel_ms, int64_t* n_launches) {
  aslsd_ctx* ctx = const_cast<aslsd_ctx*>(ctx_);
  if (!ctx) return ASLSD_ERR_ARG;
  if (ctx->timing_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ev1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess)
      ctx->last_ms = ms;
    ctx->timing_pending = false;
  }
  if (kernel_ms) *kernel_ms = ctx->last_ms;
  if (n_launches) *n_launches = ctx->last_launches;
  return ASLSD_OK;
}

// tab[g] = lower_bound(t, t0 + g*w) for g = 0..n_buckets (tab[n_buckets] = n)
__global__ void build_bucket_table_kernel(const double* __restrict__ t, int64_t n, int32_t* tab,
                                          int64_t n_buckets, double t0, double w) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_buckets) return;
  if (g == n_buckets) { tab[g] = (int32_t)n; return; }
  double x = t0 + (double)g * w;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (t[mid] < x) lo = mid + 1; else hi = mid;
  }
  tab[g] = (int32_t)lo;
}

// flag |= 1 if t is not non-decreasing (NaNs count as violations)
__global__ void check_sorted_kernel(const double* __restrict__ t, int64_t n, int* flag) {
  bool bad = false;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < n; m += (int64_t)gridDim.x * blockDim.x)
    bad |= !(t[m] >= t[m > 0 ? m - 1 : 0]);
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

int events_verdict(aslsd_ctx* ctx, bool wait) {
  if (!ctx->verdict_pending) return ASLSD_OK;
  if (!wait) {
    const cudaError_t q = cudaEventQuery(ctx->ev_flag);
    if (q == cudaErrorNotReady) return ASLSD_OK;
    if (q != cudaSuccess) return aslsd_fail(ctx, ASLSD_ERR_CUDA, "set_events: order check failed to run");
  } else {
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_flag));
  }
  ctx->verdict_pending = false;
  if (*ctx->h_flag == 0) return ASLSD_OK;
  cudaStreamSynchronize(ctx->stream);
  ctx->d = 0;
  ctx->events_epoch++;
  return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_events: times must be sorted");
}

extern "C" int aslsd_set_events(aslsd_ctx* ctx, int d, const int64_t* n_events,
                                const double* const* times, double T_f) {
  if (!ctx || d <= 0 || !n_events || !times) return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_events: bad arguments");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  for (int i = 0; i < d; ++i) {
    if (n_events[i] < 0 || n_events[i] >= (int64_t)2147483000)
      return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_events: event count out of range");
    if (n_events[i] > 0 && !times[i])
      return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_events: NULL time array of a type that has events");
  }
  // a model of another dimension cannot describe these events: it is dropped (its kernel table
  // is indexed by k * d + i everywhere), the next sum call asks for aslsd_set_model
  if (ctx->has_model && (int)ctx->kernels.size() != d * d) {
    ctx->has_model = false;
    ctx->kernels.clear();
    ctx->model_epoch++;
  }
  // layout: every type starts on a 16 B boundary and is followed by >= 2 pad doubles
  // so that aligned bulk copies may over-read
  std::vector<int64_t> off(d + 1, 0), tab_off(d + 1, 0), tab_len(d, 1);
  for (int i = 0; i < d; ++i) {
    off[i + 1] = off[i] + ((n_events[i] + 1) / 2) * 2 + 2;
    tab_len[i] = std::max<int64_t>(1, n_events[i] >> ASLSD_BUCKET_SHIFT);
    tab_off[i + 1] = tab_off[i] + tab_len[i] + 1;
  }
  const size_t n_tot = (size_t)off[d];
  // the pinned staging block may still feed the copies of the previous call
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->verdict_pending = false;    // the events it was about are replaced
  // Arrays that are already page-locked (cudaHostAlloc / cudaHostRegister, e.g. pinned torch
  // tensors) are uploaded from where they are and read in place by the host-side planning: no
  // staging copy, and their order is checked on the device behind the transfer (the host never
  // streams through them).  They must stay valid and unchanged until the next aslsd_set_events.
  std::vector<char> in_place(d, 0);
  size_t n_staged = 0;
  for (int i = 0; i < d; ++i) {
    cudaPointerAttributes attr;
    if (n_events[i] > 0 && times[i] && cudaPointerGetAttributes(&attr, times[i]) == cudaSuccess &&
        attr.type == cudaMemoryTypeHost)
      in_place[i] = 1;
    else
      n_staged += (size_t)(off[i + 1] - off[i]);
  }
  cudaGetLastError();   // older runtimes flag unregistered host memory as an error
  // one pass over the other arrays: copy into the pinned block, check the order
  if (n_staged > 0 && n_tot > ctx->h_events_cap) {
    if (ctx->h_events) CUDA_TRY(ctx, cudaFreeHost(ctx->h_events));
    ctx->h_events = nullptr;
    ctx->h_events_cap = 0;
    CUDA_TRY(ctx, cudaMallocHost((void**)&ctx->h_events, sizeof(double) * (n_tot + n_tot / 8)));
    ctx->h_events_cap = n_tot + n_tot / 8;
  }
  if (!ctx->h_pad) {
    CUDA_TRY(ctx, cudaMallocHost((void**)&ctx->h_pad, sizeof(double) * 4));
    for (int q = 0; q < 4; ++q) ctx->h_pad[q] = 1.7e308;
  }
  int rc;
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_times, &ctx->times_cap, sizeof(double) * n_tot))) return rc;
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_tab, &ctx->tab_cap, sizeof(int32_t) * (size_t)tab_off[d])))
    return rc;
  if ((rc = ensure_device_buffer(ctx, &ctx->d_dims, &ctx->dims_cap, sizeof(DimDev) * d))) return rc;
  ctx->events_epoch++;
  ctx->weight_off.clear();     // weight vectors belong to the previous events
  ctx->weight_dim.clear();
  if (d != ctx->d) ctx->model_epoch++;
  ctx->h_times.assign(d, nullptr);
  bool any_in_place = false;
  for (int i = 0; i < d; ++i) any_in_place |= in_place[i] != 0;
  // Every array page-locked: the uploads, the order checks and the bucket tables go to the copy stream,
  // one event per type, and the compute stream is ordered behind them only when it is next used
  // (CtxStream) - or, for the moment kernels of an exact-sum call, type by type as the data arrives
  // (exact.cu).  Otherwise everything stays on the compute stream as before.
  const bool on_copy_stream = n_staged == 0 && any_in_place;
  cudaStream_t up = on_copy_stream ? ctx->copy_stream : ctx->stream.raw();
  ctx->uploads_pending = false;
  if (on_copy_stream) {
    while ((int)ctx->ev_type.size() < d) {
      cudaEvent_t ev;
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      ctx->ev_type.push_back(ev);
    }
  }
  // descriptors of the types (bucket-table geometry from the first and last time of each)
  ctx->tab_t0.assign(d, 0.);
  ctx->tab_w.assign(d, 1.);
  ctx->h_dims.assign(sizeof(DimDev) * (size_t)d, 0);
  DimDev* dims0 = reinterpret_cast<DimDev*>(ctx->h_dims.data());
  for (int i = 0; i < d; ++i) {
    if (n_events[i] > 0) {
      const double t0 = times[i][0], t1 = times[i][n_events[i] - 1];
      double w = (t1 - t0) / (double)tab_len[i];
      w = w * (1. + 1e-9) + 1e-300;
      if (!(w > 0.)) w = 1.;
      ctx->tab_t0[i] = t0;
      ctx->tab_w[i] = w;
    }
    dims0[i].t = ctx->d_times + off[i];
    dims0[i].tab = ctx->d_tab + tab_off[i];
    dims0[i].n = n_events[i];
    dims0[i].n_buckets = tab_len[i];
    dims0[i].t0 = ctx->tab_t0[i];
    dims0[i].w = ctx->tab_w[i];
    dims0[i].inv_w = 1. / ctx->tab_w[i];
    dims0[i].bad = ctx->d_flag;
  }
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), up));
  if (on_copy_stream)
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_dims, dims0, sizeof(DimDev) * d, cudaMemcpyHostToDevice, up));
  // Host pass (copy into the pinned block + order check) split into segments over a few
  // threads; the thread that completes a type issues its host -> device copy, so the
  // transfer of one type overlaps the host pass over the others.
  struct Seg { int dim; int64_t a, b; };
  std::vector<Seg> segs;
  const int64_t seg_len = 1 << 19;
  std::vector<std::atomic<int>> left(d);
  for (int i = 0; i < d; ++i) {
    const int64_t n = n_events[i];
    int n_seg = 0;
    for (int64_t a = 0; !in_place[i] && (a < n || (a == 0 && n_seg == 0)); a += seg_len) {
      segs.push_back(Seg{i, a, std::min<int64_t>(n, a + seg_len)});
      ++n_seg;
    }
    left[i].store(n_seg);
    ctx->h_times[i] = in_place[i] ? times[i] : ctx->h_events + off[i];
    if (in_place[i]) {
      // the transfer starts now and runs while the threads below check the order
      const int64_t len = off[i + 1] - off[i];
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_times + off[i], times[i], sizeof(double) * (size_t)n,
                                    cudaMemcpyHostToDevice, up));
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_times + off[i] + n, ctx->h_pad, sizeof(double) * (size_t)(len - n),
                                    cudaMemcpyHostToDevice, up));
      check_sorted_kernel<<<(unsigned)std::min<int64_t>(1184, (n + 255) / 256), 256, 0, up>>>(
          ctx->d_times + off[i], n, ctx->d_flag);
      if (on_copy_stream) {
        // the type is complete once its bucket table is built: one event per type, in upload order
        // (the copy stream is in order: the event of type i covers the types before it and the
        // descriptors above)
        const int64_t nb = tab_len[i];
        build_bucket_table_kernel<<<(int)((nb + 1 + 255) / 256), 256, 0, up>>>(
            dims0[i].t, dims0[i].n, ctx->d_tab + tab_off[i], nb, dims0[i].t0, dims0[i].w);
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_type[(size_t)i], up));
      }
    }
  }
  std::atomic<size_t> next(0);
  std::atomic<int> bad(0), cuda_err(0);
  auto worker = [&]() {
    cudaSetDevice(ctx->device);
    for (;;) {
      const size_t q = next.fetch_add(1);
      if (q >= segs.size()) break;
      const Seg sg = segs[q];
      const bool stage = !in_place[sg.dim];
      double* dst = stage ? ctx->h_events + off[sg.dim] : nullptr;
      const double* src = times[sg.dim];
      bool ok = true;
      // cache-sized chunks: a branch-free (vectorisable) order check, then a plain memcpy of the
      // chunk while it is still in cache.  v >= prev is false for NaNs on either side.
      const int64_t chunk = 4096;
      for (int64_t c = sg.a; c < sg.b; c += chunk) {
        const int64_t e = std::min<int64_t>(sg.b, c + chunk);
        int okc = src[c] >= (c > 0 ? src[c - 1] : src[0]);
        for (int64_t m = c + 1; m < e; ++m) okc &= (int)(src[m] >= src[m - 1]);
        ok &= okc != 0;
        if (stage) std::memcpy(dst + c, src + c, sizeof(double) * (size_t)(e - c));
      }
      if (!ok) bad.store(1);
      if (left[sg.dim].fetch_sub(1) == 1 && stage) {
        // last segment of this type: pad with a huge finite value (the device side may
        // over-read, never uses it) and start the copy
        const int64_t len = off[sg.dim + 1] - off[sg.dim];
        for (int64_t m = n_events[sg.dim]; m < len; ++m) dst[m] = 1.7e308;
        if (cudaMemcpyAsync(ctx->d_times + off[sg.dim], dst, sizeof(double) * (size_t)len,
                            cudaMemcpyHostToDevice, up) != cudaSuccess)
          cuda_err.store(1);
      }
    }
  };
  {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    size_t cap = ASLSD_STAGE_THREADS;
    if (const char* e = std::getenv("ASLSD_STAGE_THREADS")) cap = std::max(1, std::atoi(e));
    const size_t n_thr = std::min<size_t>({(size_t)hw, cap, std::max<size_t>(1, segs.size())});
    std::vector<std::thread> pool;
    for (size_t q = 1; q < n_thr; ++q) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
  }
  if (cuda_err.load()) return aslsd_fail(ctx, ASLSD_ERR_CUDA, "set_events: host -> device copy failed");
  const bool sorted = bad.load() == 0;
  if (!sorted) {
    cudaStreamSynchronize(ctx->stream);
    ctx->d = 0;
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_events: times must be sorted");
  }
  ctx->d = d;
  ctx->T_f = T_f;
  ctx->n_events.assign(n_events, n_events + d);
  ctx->off = off;
  ctx->tab_off = tab_off;
  ctx->tab_len = tab_len;
  ctx->times_len = off[d];
  DimDev* dims = reinterpret_cast<DimDev*>(ctx->h_dims.data());
  if (!on_copy_stream) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_dims, dims, sizeof(DimDev) * d, cudaMemcpyHostToDevice, up));
    for (int i = 0; i < d; ++i) {
      const int64_t nb = tab_len[i];
      const int threads = 256;
      const int blocks = (int)((nb + 1 + threads - 1) / threads);
      build_bucket_table_kernel<<<blocks, threads, 0, up>>>(
          dims[i].t, dims[i].n, ctx->d_tab + tab_off[i], nb, dims[i].t0, dims[i].w);
    }
  }
  CUDA_TRY(ctx, cudaGetLastError());
  if (on_copy_stream) {
    ctx->n_upload_types = d;
    ctx->uploads_pending = true;
    ctx->upload_bytes = sizeof(double) * n_tot;
  }
  // no synchronisation here: the copies and the table kernels are ordered before whatever the
  // caller enqueues next on the compute stream, and they overlap its host-side planning
  if (any_in_place) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, up));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_flag, up));
    ctx->verdict_pending = true;
    // ASLSD_OPT_ASYNC_EVENTS off: wait for the device's verdict, so that an order violation is
    // reported by this call like for pageable inputs; on: the next call that waits for the
    // device reports it
    if (!ctx->async_events) return events_verdict(ctx, true);
  }
  return ASLSD_OK;
}

extern "C" int aslsd_set_horizon(aslsd_ctx* ctx, double T_f) {
  if (!ctx || !(T_f > 0.) || !std::isfinite(T_f)) return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_horizon: bad arguments");
  if (ctx->d <= 0) return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_horizon: no events set");
  if (T_f != ctx->T_f) {
    ctx->T_f = T_f;
    ctx->events_epoch++;   // interior / boundary split of every plan, cached quadrature weights
  }
  return ASLSD_OK;
}

// Per-event weight vectors: the impact values I_ki(xi^i_n) of a marked model and their
// parameter derivatives, evaluated by the reference's own impact objects on the host
// (functionals/impact_functions/impact_function.py:116-131) once per evaluation point.
extern "C" int aslsd_set_weights(aslsd_ctx* ctx, int n_vec, const int32_t* dims, const double* const* vals) {
  if (!ctx || n_vec < 0 || (n_vec > 0 && (!dims || !vals)))
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_weights: bad arguments");
  if (ctx->d == 0 || !ctx->d_times) return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_weights: no events set");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (int vrc = events_verdict(ctx, true)) return vrc;
  std::vector<int64_t> off(n_vec + 1, 0);
  for (int v = 0; v < n_vec; ++v) {
    if (dims[v] < 0 || dims[v] >= ctx->d || !vals[v])
      return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_weights: bad event type");
    off[v + 1] = off[v] + ctx->n_events[dims[v]];
  }
  int rc;
  const size_t bytes = std::max<size_t>(1, (size_t)off[n_vec]) * sizeof(double);
  if ((rc = ensure_device_buffer(ctx, (void**)&ctx->d_weights, &ctx->weights_cap, bytes))) return rc;
  if ((rc = ensure_pinned_buffer(ctx, bytes))) return rc;
  double* hp = (double*)ctx->h_pinned;
  for (int v = 0; v < n_vec; ++v)
    std::memcpy(hp + off[v], vals[v], sizeof(double) * (size_t)(off[v + 1] - off[v]));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_weights, hp, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // the pinned block is reused
  ctx->weight_off.assign(off.begin(), off.end() - 1);
  ctx->weight_dim.assign(dims, dims + n_vec);
  return ASLSD_OK;
}

extern "C" int aslsd_set_model(aslsd_ctx* ctx, int d, const int32_t* n_basis, const aslsd_basis* bases) {
  if (!ctx || d <= 0 || !n_basis || !bases) return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_model: bad arguments");
  if (ctx->d != 0 && ctx->d != d)
    return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_model: dimension differs from the events'");
  std::vector<KernelHost> ks(d * d);
  const aslsd_basis* b = bases;
  for (int k = 0; k < d; ++k)
    for (int i = 0; i < d; ++i) {
      KernelHost& K = ks[k * d + i];
      int nb = n_basis[k * d + i];
      if (nb <= 0) return aslsd_fail(ctx, ASLSD_ERR_ARG, "set_model: kernel without bases");
      for (int r = 0; r < nb; ++r, ++b) {
        if (b->type != ASLSD_BASIS_EXPONENTIAL && b->type != ASLSD_BASIS_GAUSSIAN)
          return aslsd_fail(ctx, ASLSD_ERR_UNSUPPORTED,
                            "set_model: only ExponentialKernel and GaussianKernel bases are supported");
        BasisHost h;
        h.type = b->type;
        int nv = (b->type == ASLSD_BASIS_EXPONENTIAL) ? 2 : 3;
        h.free_mask = b->free_mask & ((1 << nv) - 1);
        for (int v = 0; v < 3; ++v) h.fixed[v] = b->fixed[v];
        K.bases.push_back(h);
      }
    }
  // Upsilon closures of the reference are all bound to the kernel objects of
  // the last loop iteration (mhp.py:527-536): K[d-1][d-1] provides the basis
  // structure of every upsilon term, so all kernels must share it.
  const KernelHost& KU = ks[d * d - 1];
  for (auto& K : ks) {
    if (K.bases.size() != KU.bases.size())
      return aslsd_fail(ctx, ASLSD_ERR_UNSUPPORTED,
                        "set_model: all kernels must have the same basis structure");
    for (size_t r = 0; r < K.bases.size(); ++r) {
      if (K.bases[r].type != KU.bases[r].type || K.bases[r].free_mask != KU.bases[r].free_mask)
        return aslsd_fail(ctx, ASLSD_ERR_UNSUPPORTED,
                          "set_model: all kernels must have the same basis structure");
    }
  }
  ctx->kernels.swap(ks);
  ctx->has_model = true;
  ctx->model_epoch++;
  if (ctx->d == 0) ctx->d = d;
  return ASLSD_OK;
}

extern "C" int aslsd_last_moment_timing(const aslsd_ctx* ctx_, double* kernel_ms) {
  aslsd_ctx* ctx = const_cast<aslsd_ctx*>(ctx_);
  if (!ctx || !kernel_ms) return ASLSD_ERR_ARG;
  if (ctx->moment_timing_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ev_q1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ctx->ev_q0, ctx->ev_q1) == cudaSuccess)
      ctx->last_moment_ms = ms;
    ctx->moment_timing_pending = false;
  }
  *kernel_ms = ctx->last_moment_ms;
  return ASLSD_OK;
}

extern "C" int aslsd_model_n_param(const aslsd_ctx* ctx, int k, int32_t* n_param_k, int32_t* n_param) {
  if (!ctx || !ctx->has_model || k < 0 || k >= ctx->d) return ASLSD_ERR_ARG;
  int tot = 1;
  for (int i = 0; i < ctx->d; ++i) {
    int n = ctx->kernels[k * ctx->d + i].n_param();
    if (n_param) n_param[i] = n;
    tot += n;
  }
  if (n_param_k) *n_param_k = tot;
  return ASLSD_OK;
}

// ---- fp64 roofline denominator -----------------------------------------------------------------
// MEASURED_PEAKS.json has no fp64 entry: a DFMA microbenchmark
// with 8 independent chains per thread, enough resident warps to saturate the
// fp64 pipe, timed with CUDA events on the context's stream.
__global__ void __launch_bounds__(256)
dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3.;
  double x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[0] = s;   // keep the chains alive
}

extern "C" int aslsd_measure_fp64_peak(aslsd_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return ASLSD_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  double* d_out = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&d_out, sizeof(double)));
  const int iters = 1 << 14, blocks = 148 * 8, threads = 256;
  double best = 0.;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(ctx->ev0, ctx->stream);
    dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 0.999999, 1e-9);
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaError_t e = cudaEventSynchronize(ctx->ev1);
    if (e != cudaSuccess) { cudaFree(d_out); CUDA_TRY(ctx, e); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    best = std::max(best, flops / (ms * 1e-3) * 1e-12);
  }
  cudaFree(d_out);
  *tflops = best;
  return ASLSD_OK;
}
