// K5 for large state spaces -- placeholder until the DMMA downward kernel lands.
#include "rt_common.cuh"
int rt_posterior_dmma_dispatch(int, int, int64_t, int64_t, const int32_t*, const int32_t*, int,
                               const double*, const double*, const void*, const double*,
                               const int8_t*, double*, double*, double*, cudaStream_t) {
  return RT_ERR_UNSUPPORTED;
}
