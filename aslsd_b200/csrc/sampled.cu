// Sampled-lag sums of the stratified (ASLSD) estimator -- placeholder until the
// kernels land; see include/aslsd_b200.h.
#include "common.cuh"

extern "C" int aslsd_sampled_sums(aslsd_ctx* ctx, int, const double*, int, int, int, int, int, const int64_t*,
                                  const int64_t*, const int64_t*, double*, double*, double*, double*) {
  return aslsd_fail(ctx, ASLSD_ERR_UNSUPPORTED, "sampled sums: not built yet");
}
extern "C" int aslsd_update_allocation(aslsd_ctx* ctx, int, const double*, const double*, const double*,
                                       double*) {
  return aslsd_fail(ctx, ASLSD_ERR_UNSUPPORTED, "allocation update: not built yet");
}
