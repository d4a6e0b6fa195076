// TEST INFRASTRUCTURE ONLY - a do-nothing CUDA runtime for running the HOST side of
// libaslsd_b200.so (argument checks, planning, packing, assembly, memory management) on a machine
// without a GPU, under AddressSanitizer (tests/test_host_code_sanitizer.py).
//
// It is not a CPU path of the product and cannot serve as one: kernel launches are no-ops, so
// no sum, index map or statistic is ever computed - "device" buffers are zero-filled host
// allocations whose only purpose is that the sanitizer sees every copy the host code makes
// into and out of them.  The product library (aslsd_b200/_lib) links the real runtime and is
// never loaded together with this file.
#include <cuda_runtime_api.h>

#include <cstdlib>
#include <cstring>

extern "C" {

// ---- registration hooks emitted by nvcc for every translation unit with device code ----------
void** __cudaRegisterFatBinary(void*) {
  static void* handle = nullptr;
  return &handle;
}
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char*, char*, const char*, int, uint3*, uint3*, dim3*, dim3*, int*) {}
void __cudaRegisterVar(void**, char*, char*, const char*, int, size_t, int, int) {}

static thread_local struct { dim3 grid, block; size_t smem; void* stream; } g_cfg;
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, struct CUstream_st* stream) {
  g_cfg.grid = grid; g_cfg.block = block; g_cfg.smem = smem; g_cfg.stream = stream;
  return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3* grid, dim3* block, size_t* smem, void* stream) {
  *grid = g_cfg.grid; *block = g_cfg.block; *smem = g_cfg.smem; *(void**)stream = g_cfg.stream;
  return cudaSuccess;
}

// ---- launches: nothing runs ---------------------------------------------------------------------
static long g_launches = 0;
cudaError_t cudaLaunchKernel(const void*, dim3 grid, dim3 block, void**, size_t, cudaStream_t) {
  // a launch geometry no device accepts is still an error worth seeing
  if (grid.x == 0 || grid.y == 0 || grid.z == 0 || block.x * block.y * block.z == 0 ||
      block.x * block.y * block.z > 1024)
    return cudaErrorInvalidConfiguration;
  ++g_launches;
  return cudaSuccess;
}
long aslsd_stub_launch_count() { return g_launches; }

// ---- memory: plain host allocations ----------------------------------------------------------------
cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
// One-shot substitution of a device -> host copy: the next such copy of exactly `bytes` bytes takes
// its data from `src`.  The plan / assembly checks use it to hand rows computed with numpy to the
// fused entry points (aslsd_exact_sums_weighted), whose kernels ran as no-ops.
static const void* g_next_src = nullptr;
static size_t g_next_bytes = 0;
void aslsd_stub_next_d2h(const void* src, size_t bytes) { g_next_src = src; g_next_bytes = bytes; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind kind, cudaStream_t) {
  if (kind == cudaMemcpyDeviceToHost && g_next_src && n == g_next_bytes) {
    std::memcpy(dst, g_next_src, n);
    g_next_src = nullptr;
    return cudaSuccess;
  }
  std::memcpy(dst, src, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpyToSymbol(const void*, const void*, size_t, size_t, cudaMemcpyKind) { return cudaSuccess; }
cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) {
  std::memset(a, 0, sizeof(*a));
  a->type = cudaMemoryTypeUnregistered;   // every caller buffer is pageable host memory
  return cudaSuccess;
}

// ---- device, streams, events ---------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int dev) { return dev == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "stub runtime error"; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)std::malloc(1); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)std::malloc(1); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)std::malloc(1); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

// what the CUB dispatch layers (simulate.cu: scan, radix sort) ask the runtime before launching
cudaError_t cudaGetDevice(int* dev) { *dev = 0; return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  *v = a == cudaDevAttrMultiProcessorCount ? 148 : a == cudaDevAttrComputeCapabilityMajor ? 10
     : a == cudaDevAttrComputeCapabilityMinor ? 0 : 65536;
  return cudaSuccess;
}
cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, const void*) {
  std::memset(a, 0, sizeof(*a));
  a->maxThreadsPerBlock = 1024;
  a->ptxVersion = a->binaryVersion = 100;
  return cudaSuccess;
}
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n, const void*, int, size_t, unsigned) {
  *n = 1;
  return cudaSuccess;
}

}  // extern "C"
