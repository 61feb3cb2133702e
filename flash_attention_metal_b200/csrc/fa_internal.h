// Internal declarations shared by the translation units of libflash_attn_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>

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

// "once per CUDA device" guard for per-device function attributes (max dynamic shared memory).
// Thread-safe (one host thread per GPU is a supported calling pattern); a device is marked done only
// after its set-up succeeded, so a transient failure is retried by the next call.
struct DeviceOnce {
  std::atomic<bool> done[64];
  std::mutex mu;
  DeviceOnce() { for (auto &d : done) d.store(false, std::memory_order_relaxed); }
  template <typename F>
  int run(F &&setup) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return setup();
    if (done[dev].load(std::memory_order_acquire)) return FA_OK;
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev].load(std::memory_order_relaxed)) return FA_OK;
    const int rc = setup();
    if (rc == FA_OK) done[dev].store(true, std::memory_order_release);
    return rc;
  }
};

// L2 budget (MB) of one dispatch group of heads (sched.cuh).  48 is the measured optimum on B200's
// 126 MB L2; only tests change it (fa_debug_set_l2_group_mb) to prove results do not depend on it.
int l2_group_mb();

// Load every tensor-core / ring kernel into the current device's context and apply the per-device function
// attributes.  With CUDA's lazy module loading a kernel is otherwise loaded at its first launch, which
// synchronises the context -- fatal once a stream of that context is parked on a flag that a later enqueue has
// yet to write (single-process multi-GPU ring).  fa_mgpu_create runs it for every device of the group.
int preload_fwd_tc();
int preload_bwd_tc();
int preload_bwd_fused();
int preload_ring();
int preload_kernels();

// SM count of the current device (cached per device, thread-safe; 148 if the query fails)
int device_sm_count();

// Restores the caller's current device when an entry point that has to switch devices returns.
struct DeviceGuard {
  int saved = -1;
  DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) { saved = -1; (void)cudaGetLastError(); } }
  ~DeviceGuard() { if (saved >= 0) (void)cudaSetDevice(saved); }
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

// Ring attention: fold this launch's partial result into a running fp32 (O_acc, L_acc) pair in the
// kernel epilogue (modes per row range: rows < half_rows use `lo`, the rest `hi`).
//   0 none: plain launch      1 first: acc = partial      2 middle: acc = acc (+) partial
//   3 last : O, L = acc (+) partial (16-bit O, final L)
struct FwdMerge {
  float *O_acc;   // fp32, addressed like O (same strides)
  float *L_acc;   // fp32, addressed like L
  int lo, hi, half_rows;
};

// rectangular (Nq x Nk, separate Q/O and K/V strides) form used by ring attention
int launch_fwd_tc_rect(const void *Q, const void *K, const void *V, void *O, float *L, int Nq, int Nk,
                       int D, float scale, int64_t q_batch_stride, int64_t q_head_stride,
                       int64_t kv_batch_stride, int64_t kv_head_stride, int is_causal, int B, int H,
                       int dtype, cudaStream_t stream, const FwdMerge *merge = nullptr);
void set_fwd_split_max(int cap);  // development aid behind fa_debug_set_fwd_split_max

#if defined(FA_FWD_TRACE) || defined(FA_BWD_TRACE)
extern long long *g_trace_buffer;  // trace builds only (never in the product build): see fa_debug_set_prof_buffer
#endif

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
                       cudaStream_t stream, void *fused_sems = nullptr);

// fused five-GEMM backward (bwd_fused.cu): `sems` = bwd_fused_sem_bytes() bytes of ordering counters
size_t bwd_fused_sem_bytes(int Nq, int B, int H);
int launch_bwd_fused(const void *Q, const void *K, const void *V, const void *dO, const float *L, const float *delta,
                     float *dQ, float *dK, float *dV, int Nq, int Nk, int D, float scale, int64_t q_batch_stride,
                     int64_t q_head_stride, int64_t kv_batch_stride, int64_t kv_head_stride, int is_causal, int acc_dq,
                     int B, int H, int dtype, void *sems, cudaStream_t stream);
// FA_BWD_TWO_KERNEL (default) or FA_BWD_FUSED: fa_set_backward_algorithm
int bwd_mode();


}  // namespace fa
