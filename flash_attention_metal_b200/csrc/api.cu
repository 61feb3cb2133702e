// extern "C" entry points of libflash_attn_b200.so (see include/flash_attn_b200.h).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "fa_internal.h"
#include "sched.cuh"

namespace fa {

static thread_local char g_err[512] = "";
static thread_local long g_launches = 0;

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches += n; }

}  // namespace fa

#if defined(FA_FWD_TRACE) || defined(FA_BWD_TRACE)
namespace fa { long long *g_trace_buffer = nullptr; }
#endif

namespace fa {
static std::atomic<int> g_bwd_mode{0};
int bwd_mode() { return g_bwd_mode.load(std::memory_order_relaxed); }

static std::atomic<int> g_l2_group_mb{48};
int l2_group_mb() { return g_l2_group_mb.load(std::memory_order_relaxed); }

int preload_kernels() {
  int rc;
  if ((rc = preload_fwd_tc()) || (rc = preload_bwd_tc()) || (rc = preload_bwd_fused()) || (rc = preload_ring())) return rc;
  return FA_OK;
}

int device_sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
}  // namespace fa

using namespace fa;

extern "C" {

const char *fa_last_error(void) { return g_err; }
int fa_version(void) { return 100; }
long fa_launch_count(void) { return g_launches; }
void fa_reset_launch_count(void) { g_launches = 0; }

#if defined(FA_FWD_TRACE) || defined(FA_BWD_TRACE)
// Trace builds only (-DFA_FWD_TRACE / -DFA_BWD_TRACE, tests/trace_probe*.py): device buffer the kernels
// fill with clock64() timestamps of one CTA; NULL switches it off.  Not compiled into the product.
void fa_debug_set_prof_buffer(long long *dev) { g_trace_buffer = dev; }
#endif

// Development aid (not in the public header): one ring-attention partial on one GPU -- the forward over
// a chunk of keys whose epilogue folds the result into the running fp32 (O_acc, L_acc) pair (FwdMerge:
// lo / hi = 0 none, 1 first, 2 middle, 3 last, for rows below / from half_rows).  Tests chain it over
// key chunks to check the fused merge without a second GPU.
int fa_debug_forward_partial(const void *Q, const void *K, const void *V, void *O, float *L, float *O_acc, float *L_acc,
                             int Nq, int Nk, int D, float scale, int H, int is_causal, int lo, int hi, int half_rows,
                             int dtype, fa_stream_t stream) {
  FwdMerge m{O_acc, L_acc, lo, hi, half_rows};
  return launch_fwd_tc_rect(Q, K, V, O, L, Nq, Nk, D, scale, (int64_t)H * Nq * D, (int64_t)Nq * D, (int64_t)H * Nk * D,
                            (int64_t)Nk * D, is_causal, 1, H, dtype, (cudaStream_t)stream, &m);
}

int fa_set_backward_algorithm(int algorithm) {
  FA_REQUIRE(algorithm == FA_BWD_TWO_KERNEL || algorithm == FA_BWD_FUSED, "unknown backward algorithm %d", algorithm);
  g_bwd_mode.store(algorithm, std::memory_order_relaxed);
  return FA_OK;
}
int fa_get_backward_algorithm(void) { return g_bwd_mode.load(std::memory_order_relaxed); }

// Development aid (not in the public header): L2 budget of a dispatch group of heads, in MB.
void fa_debug_set_l2_group_mb(int mb) { g_l2_group_mb.store(mb < 0 ? 0 : mb, std::memory_order_relaxed); }

// Development aid (not in the public header): cap of the key-split cluster size of small forward
// launches (default 4; tests exercise 8).
void fa_debug_set_fwd_split_max(int cap) { set_fwd_split_max(cap); }

// Development aid (not in the public header): the dispatch geometry sched.cuh picks for a launch
// of `n_blocks` blocks per head -- out = {heads per group, grid.x, grid.y, grid.z}.  Host logic only.
void fa_debug_dispatch(int uneven_work, long long streamed_bytes_per_head, int n_blocks, int H, int B, int *out) {
  const int group = dispatch_group(uneven_work != 0, streamed_bytes_per_head, H * B);
  const dim3 g = dispatch_grid(group, n_blocks, H, B);
  out[0] = group; out[1] = (int)g.x; out[2] = (int)g.y; out[3] = (int)g.z;
}

int fa_preload_kernels(void) { return preload_kernels(); }

int fa_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return n;
}

int naive_attention(const float *Q, const float *K, const float *V, float *O, int N, int D,
                    float scale, int is_causal, fa_stream_t stream) {
  return launch_fp32(0, Q, K, V, O, N, D, scale, 0, 0, is_causal, 1, 1, (cudaStream_t)stream);
}

int flash_attention(const float *Q, const float *K, const float *V, float *O, int N, int D,
                    float scale, int is_causal, fa_stream_t stream) {
  return launch_fp32(1, Q, K, V, O, N, D, scale, 0, 0, is_causal, 1, 1, (cudaStream_t)stream);
}

int flash_attention_v2(const float *Q, const float *K, const float *V, float *O, int N, int D,
                       float scale, int is_causal, fa_stream_t stream) {
  return launch_fp32(2, Q, K, V, O, N, D, scale, 0, 0, is_causal, 1, 1, (cudaStream_t)stream);
}

int flash_attention_v2_batched(const float *Q, const float *K, const float *V, float *O, int N,
                               int D, float scale, int64_t batch_stride, int64_t head_stride,
                               int is_causal, int B, int H, fa_stream_t stream) {
  return launch_fp32(2, Q, K, V, O, N, D, scale, batch_stride, head_stride, is_causal, B, H,
                     (cudaStream_t)stream);
}

int flash_attention_simd(const void *Q, const void *K, const void *V, void *O, int N, int D,
                         float scale, int dtype, fa_stream_t stream) {
  return launch_fwd_tc(Q, K, V, O, nullptr, N, D, scale, (int64_t)N * D, (int64_t)N * D, 0, 1, 1,
                       dtype, (cudaStream_t)stream);
}

int flash_attention_v4_half(const void *Q, const void *K, const void *V, void *O, int N, int D,
                            float scale, int64_t batch_stride, int64_t head_stride, float *L_out,
                            int is_causal, int B, int H, int dtype, fa_stream_t stream) {
  return launch_fwd_tc(Q, K, V, O, L_out, N, D, scale, batch_stride, head_stride, is_causal, B, H,
                       dtype, (cudaStream_t)stream);
}

int flash_attention_v4_half_rect(const void *Q, const void *K, const void *V, void *O, int Nq, int Nk,
                                 int D, float scale, int64_t q_batch_stride, int64_t q_head_stride,
                                 int64_t kv_batch_stride, int64_t kv_head_stride, float *L_out, int B,
                                 int H, int dtype, fa_stream_t stream) {
  return launch_fwd_tc_rect(Q, K, V, O, L_out, Nq, Nk, D, scale, q_batch_stride, q_head_stride, kv_batch_stride,
                            kv_head_stride, 0, B, H, dtype, (cudaStream_t)stream);
}

int flash_attention_backward(const void *Q, const void *K, const void *V, const void *O,
                             const void *dO, const float *L, float *dQ, float *dK, float *dV,
                             int N, int D, float scale, int64_t batch_stride,
                             int64_t head_stride, int is_causal, int B, int H, int dtype,
                             void *workspace, size_t workspace_bytes, fa_stream_t stream) {
  return launch_bwd_tc(Q, K, V, O, dO, L, dQ, dK, dV, N, D, scale, batch_stride, head_stride,
                       is_causal, B, H, dtype, workspace, workspace_bytes, (cudaStream_t)stream);
}

int flash_attention_backward_rect(const void *Q, const void *K, const void *V, const void *dO, const float *L,
                                  const float *delta, float *dQ, float *dK, float *dV, int Nq, int Nk, int D,
                                  float scale, int64_t q_batch_stride, int64_t q_head_stride,
                                  int64_t kv_batch_stride, int64_t kv_head_stride, int acc_dq, int B, int H,
                                  int dtype, fa_stream_t stream) {
  FA_REQUIRE(dQ != nullptr || dK != nullptr, "nothing to compute: dQ, dK and dV are all null");
  // no workspace in this call's signature: the two-kernel form (needs no ordering counters)
  return launch_bwd_tc_rect(Q, K, V, dO, L, delta, dQ, dK, dV, Nq, Nk, D, scale, q_batch_stride, q_head_stride,
                            kv_batch_stride, kv_head_stride, 0, acc_dq, B, H, dtype, (cudaStream_t)stream, nullptr);
}

int fa_rowsum_delta(const void *O, const void *dO, float *delta, int N, int D, int64_t batch_stride,
                    int64_t head_stride, int B, int H, int dtype, fa_stream_t stream) {
  FA_REQUIRE(O && dO && delta && N >= 1 && B >= 1 && H >= 1, "bad arguments");
  FA_REQUIRE(D == 64 || D == 128, "D must be 64 or 128 (got %d)", D);
  FA_REQUIRE(dtype == FA_DTYPE_FP16 || dtype == FA_DTYPE_BF16, "dtype must be FA_DTYPE_FP16 or FA_DTYPE_BF16");
  FA_REQUIRE(aligned16(O) && aligned16(dO), "O and dO must be 16-byte aligned");
  FA_REQUIRE(batch_stride % D == 0 && head_stride % D == 0, "strides must be multiples of D");
  return launch_bwd_delta(O, dO, delta, N, D, batch_stride, head_stride, B, H, dtype, (cudaStream_t)stream);
}

size_t fa_workspace_bytes_backward(int N, int D, int B, int H) {
  (void)D;
  if (N < 1 || B < 1 || H < 1) return 0;
  // ordering counters of the fused kernel (one per head and 128-row query tile), then
  // D_i = rowsum(dO o O): one float per (b, h, i); both rounded up to 256 bytes
  size_t bytes = (size_t)B * H * N * sizeof(float);
  return bwd_fused_sem_bytes(N, B, H) + ((bytes + 255) & ~(size_t)255);
}

}  // extern "C"
