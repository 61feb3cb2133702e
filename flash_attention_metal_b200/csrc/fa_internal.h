// Internal declarations shared by the translation units of libflash_attn_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "flash_attn_b200.h"

namespace fa {

// thread-local last-error string + launch counter (api.cu)
int set_error(int code, const char *fmt, ...);
void count_launch(int n = 1);

#define FA_CUDA_CHECK(expr)                                                            \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      return ::fa::set_error(FA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                             cudaGetErrorString(_e), __FILE__, __LINE__);              \
  } while (0)

#define FA_REQUIRE(cond, ...)                                                          \
  do {                                                                                 \
    if (!(cond)) return ::fa::set_error(FA_ERR_INVALID, __VA_ARGS__);                  \
  } while (0)

// "once per CUDA device" flag for per-device function attributes (max dynamic shared memory)
struct DeviceOnce {
  bool done[64] = {};
  bool first_use() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// fp32 variants (fp32_kernels.cu).  variant: 0 naive, 1 tiled v1, 2 vectorised v2.
int launch_fp32(int variant, const float *Q, const float *K, const float *V, float *O, int N, int D,
                float scale, int64_t batch_stride, int64_t head_stride, int is_causal, int B, int H,
                cudaStream_t stream);

// tcgen05 forward (fwd_tc.cu)
int launch_fwd_tc(const void *Q, const void *K, const void *V, void *O, float *L, int N, int D,
                  float scale, int64_t batch_stride, int64_t head_stride, int is_causal, int B, int H,
                  int dtype, cudaStream_t stream);

// rectangular (Nq x Nk, separate Q/O and K/V strides) form used by ring attention
int launch_fwd_tc_rect(const void *Q, const void *K, const void *V, void *O, float *L, int Nq, int Nk,
                       int D, float scale, int64_t q_batch_stride, int64_t q_head_stride,
                       int64_t kv_batch_stride, int64_t kv_head_stride, int is_causal, int B, int H,
                       int dtype, cudaStream_t stream);

extern long long *g_fwd_prof;  // development aid: phase-timing buffer, see fa_debug_set_prof_buffer

// backward (bwd_tc.cu)
int launch_bwd_tc(const void *Q, const void *K, const void *V, const void *O, const void *dO,
                  const float *L, float *dQ, float *dK, float *dV, int N, int D, float scale,
                  int64_t batch_stride, int64_t head_stride, int is_causal, int B, int H, int dtype,
                  void *workspace, size_t workspace_bytes, cudaStream_t stream);

int launch_bwd_delta(const void *O, const void *dO, float *delta, int Nq, int D, int64_t batch_stride,
                     int64_t head_stride, int B, int H, int dtype, cudaStream_t stream);
int launch_bwd_tc_rect(const void *Q, const void *K, const void *V, const void *dO, const float *L,
                       const float *delta, float *dQ, float *dK, float *dV, int Nq, int Nk, int D, float scale,
                       int64_t q_batch_stride, int64_t q_head_stride, int64_t kv_batch_stride,
                       int64_t kv_head_stride, int is_causal, int acc_dq, int B, int H, int dtype,
                       cudaStream_t stream);

}  // namespace fa
