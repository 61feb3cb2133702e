// Backward pass (placeholder until the tcgen05 kernels land in this file).
#include "fa_internal.h"

namespace fa {

int launch_bwd_tc(const void *, const void *, const void *, const void *, const void *, const float *,
                  float *, float *, float *, int, int, float, int64_t, int64_t, int, int, int, int,
                  void *, size_t, cudaStream_t) {
  return set_error(FA_ERR_UNSUPPORTED, "flash_attention_backward is not implemented yet");
}

}  // namespace fa
