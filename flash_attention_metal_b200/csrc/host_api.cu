// Host-buffer entry points (fa_host_*): the one-call form for a caller whose tensors live
// in host memory, as the reference's do (MTLResourceStorageModeShared, main.mm:104-115).
//
// Device scratch comes from a grow-only pool per device (released by fa_host_release), so
// repeated calls do not pay cudaMalloc.  The 16-bit calls are pipelined over groups of heads
// on three streams -- host->device copies of group g+1, kernels of group g and device->host
// copies of group g-1 overlap -- because heads are independent (kernels.metal:622).  Pass
// pinned host memory (cudaHostAlloc / cudaHostRegister) for full PCIe speed; pageable memory
// works but serialises the copies.
#include <cuda_runtime.h>

#include <mutex>

#include "fa_internal.h"

namespace fa {
namespace {

struct HostPool {
  static constexpr int kSlots = 13;
  void *ptr[kSlots] = {};
  size_t cap[kSlots] = {};
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  static constexpr int kMaxGroups = 64;
  cudaEvent_t in_done[kMaxGroups] = {}, run_done[kMaxGroups] = {};
  std::mutex mu;

  // streams and events are created on first use, on the pool's own device (the caller's current device)
  int ensure_streams() {
    if (!s_in) {
      FA_CUDA_CHECK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
      FA_CUDA_CHECK(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
      FA_CUDA_CHECK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
      for (int i = 0; i < kMaxGroups; ++i) {
        FA_CUDA_CHECK(cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming));
        FA_CUDA_CHECK(cudaEventCreateWithFlags(&run_done[i], cudaEventDisableTiming));
      }
    }
    return FA_OK;
  }
  int get(int slot, size_t bytes, void **out) {
    if (cap[slot] < bytes) {
      if (ptr[slot]) cudaFree(ptr[slot]);
      ptr[slot] = nullptr;
      cap[slot] = 0;
      FA_CUDA_CHECK(cudaMalloc(&ptr[slot], bytes));
      cap[slot] = bytes;
    }
    *out = ptr[slot];
    return FA_OK;
  }
  void release() {
    for (int i = 0; i < kSlots; ++i) {
      if (ptr[i]) cudaFree(ptr[i]);
      ptr[i] = nullptr;
      cap[i] = 0;
    }
    if (s_in) {
      cudaStreamDestroy(s_in); cudaStreamDestroy(s_run); cudaStreamDestroy(s_out);
      for (int i = 0; i < kMaxGroups; ++i) { cudaEventDestroy(in_done[i]); cudaEventDestroy(run_done[i]); }
      s_in = s_run = s_out = nullptr;
    }
  }
};

// One pool per device: a process that drives several GPUs (one host thread per GPU, or one thread switching
// devices) keeps every device's scratch and streams; calls on different devices do not serialise on one lock.
constexpr int kMaxDevices = 64;
HostPool g_pools[kMaxDevices];

int current_pool(HostPool **out) {
  int dev = 0;
  FA_CUDA_CHECK(cudaGetDevice(&dev));
  FA_REQUIRE(dev >= 0 && dev < kMaxDevices, "device index %d out of range", dev);
  *out = &g_pools[dev];
  return FA_OK;
}

// error path after something has been enqueued: do not return while copies may still touch the caller's buffers
int fail_after_enqueue(HostPool &pool, int rc) {
  if (pool.s_in) { cudaStreamSynchronize(pool.s_in); cudaStreamSynchronize(pool.s_run); cudaStreamSynchronize(pool.s_out); }
  (void)cudaGetLastError();
  return rc;
}

// fp32 -> 16-bit (round to nearest even), 8 elements per thread: the optional 16-bit gradient output of
// the host-buffer call (halves the device->host bytes, which bound that call)
template <int IS_BF16>
__global__ void __launch_bounds__(256) narrow_kernel(const float *__restrict__ src, uint16_t *__restrict__ dst, size_t n8) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = reinterpret_cast<const float4 *>(src)[2 * i], b = reinterpret_cast<const float4 *>(src)[2 * i + 1];
  uint32_t w[4];
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (IS_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[k]) : "f"(v[2 * k + 1]), "f"(v[2 * k]));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w[k]) : "f"(v[2 * k + 1]), "f"(v[2 * k]));
  }
  reinterpret_cast<uint4 *>(dst)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

int host_half_impl(const void *Q, const void *K, const void *V, const void *dO, void *O, float *L,
                   void *dQ, void *dK, void *dV, int N, int D, float scale, int is_causal, int B,
                   int H, int dtype, int grad_dtype = -1) {
  const bool bwd = dO != nullptr;
  const bool narrow = grad_dtype >= 0;  // gradients leave the device as 16-bit values
  FA_REQUIRE(Q && K && V && O && N >= 1 && B >= 1 && H >= 1, "bad arguments");
  FA_REQUIRE(D == 64 || D == 128, "D must be 64 or 128 (got %d)", D);
  FA_REQUIRE(!bwd || (dQ && dK && dV), "backward needs dQ, dK and dV");
  HostPool *pool_ptr = nullptr;
  int rc = current_pool(&pool_ptr);
  if (rc != FA_OK) return rc;
  HostPool &g_pool = *pool_ptr;
  std::lock_guard<std::mutex> lock(g_pool.mu);
  rc = g_pool.ensure_streams();
  if (rc != FA_OK) return rc;
  const int heads = B * H;
  const size_t head_elems = (size_t)N * D;
  const size_t hb = head_elems * 2, hf = head_elems * 4, hl = (size_t)N * 4;
  void *dq_, *dk_, *dv_, *do_, *dOut, *dL, *gQ = nullptr, *gK = nullptr, *gV = nullptr, *dDelta = nullptr;
  if ((rc = g_pool.get(0, heads * hb, &dq_)) || (rc = g_pool.get(1, heads * hb, &dk_)) ||
      (rc = g_pool.get(2, heads * hb, &dv_)) || (rc = g_pool.get(3, heads * hb, &dOut)) ||
      (rc = g_pool.get(4, heads * hl, &dL)))
    return rc;
  do_ = nullptr;
  if (bwd) {
    if ((rc = g_pool.get(5, heads * hb, &do_)) || (rc = g_pool.get(6, heads * hf, &gQ)) ||
        (rc = g_pool.get(7, heads * hf, &gK)) || (rc = g_pool.get(8, heads * hf, &gV)) ||
        (rc = g_pool.get(9, fa_workspace_bytes_backward(N, D, 1, heads) + 256, &dDelta)))
      return rc;
  }
  void *nQ = nullptr, *nK = nullptr, *nV = nullptr;
  if (bwd && narrow &&
      ((rc = g_pool.get(10, heads * hb, &nQ)) || (rc = g_pool.get(11, heads * hb, &nK)) || (rc = g_pool.get(12, heads * hb, &nV))))
    return rc;
  // Groups of heads.  Steady state: enough CTAs per group to fill the GPU, enough groups to overlap
  // copies.  The device->host direction carries the most bytes (fp32 gradients), so the pipeline is
  // bound by how early the first result copy can start: the first groups are small (1, 1, 2 heads)
  // to get results onto the bus quickly, later ones grow to the steady size.
  int per_group = (heads + 7) / 8;
  const int min_heads = (int)((148 * 256 + N - 1) / N);  // ~one wave of 256-row CTAs
  if (per_group < min_heads) per_group = min_heads;
  if (per_group > heads) per_group = heads;
  int group_h0[HostPool::kMaxGroups], group_nh[HostPool::kMaxGroups], groups = 0;
  {
    const int ramp[3] = {1, 1, 2};
    int h0 = 0;
    while (h0 < heads) {
      const int left = heads - h0;
      int nh = (groups < 3 && ramp[groups] < per_group) ? ramp[groups] : per_group;
      const int groups_left = HostPool::kMaxGroups - groups;  // never run out of events
      if ((left + nh - 1) / nh > groups_left) nh = (left + groups_left - 1) / groups_left;
      if (nh > left) nh = left;
      group_h0[groups] = h0;
      group_nh[groups] = nh;
      ++groups;
      h0 += nh;
    }
  }
  auto at = [](const void *p, size_t off) { return (const void *)((const char *)p + off); };
  auto atw = [](void *p, size_t off) { return (void *)((char *)p + off); };
  // everything below only enqueues; on any failure the streams are drained before the error is returned, so
  // that no copy still touches the caller's buffers after the call
  auto enqueue_all = [&]() -> int {
  for (int g = 0; g < groups; ++g) {
    const int h0 = group_h0[g], nh = group_nh[g];
    const size_t ob = (size_t)h0 * hb, of = (size_t)h0 * hf, ol = (size_t)h0 * hl;
    FA_CUDA_CHECK(cudaMemcpyAsync(atw(dq_, ob), at(Q, ob), nh * hb, cudaMemcpyHostToDevice, g_pool.s_in));
    FA_CUDA_CHECK(cudaMemcpyAsync(atw(dk_, ob), at(K, ob), nh * hb, cudaMemcpyHostToDevice, g_pool.s_in));
    FA_CUDA_CHECK(cudaMemcpyAsync(atw(dv_, ob), at(V, ob), nh * hb, cudaMemcpyHostToDevice, g_pool.s_in));
    if (bwd) FA_CUDA_CHECK(cudaMemcpyAsync(atw(do_, ob), at(dO, ob), nh * hb, cudaMemcpyHostToDevice, g_pool.s_in));
    FA_CUDA_CHECK(cudaEventRecord(g_pool.in_done[g], g_pool.s_in));
    FA_CUDA_CHECK(cudaStreamWaitEvent(g_pool.s_run, g_pool.in_done[g], 0));
    rc = launch_fwd_tc(atw(dq_, ob), atw(dk_, ob), atw(dv_, ob), atw(dOut, ob), (float *)atw(dL, ol), N, D, scale,
                       (int64_t)nh * head_elems, (int64_t)head_elems, is_causal, 1, nh, dtype, g_pool.s_run);
    if (rc != FA_OK) return rc;
    if (bwd) {
      rc = launch_bwd_tc(atw(dq_, ob), atw(dk_, ob), atw(dv_, ob), atw(dOut, ob), atw(do_, ob), (float *)atw(dL, ol),
                         (float *)atw(gQ, of), (float *)atw(gK, of), (float *)atw(gV, of), N, D, scale,
                         (int64_t)nh * head_elems, (int64_t)head_elems, is_causal, 1, nh, dtype, dDelta,
                         fa_workspace_bytes_backward(N, D, 1, nh), g_pool.s_run);  // one scratch: the groups run in order
      if (rc != FA_OK) return rc;
      if (narrow) {
        const size_t n8 = (size_t)nh * head_elems / 8;
        const unsigned blocks = (unsigned)((n8 + 255) / 256);
        void *src[3] = {gQ, gK, gV}, *dst[3] = {nQ, nK, nV};
        for (int t = 0; t < 3; ++t) {
          if (grad_dtype == FA_DTYPE_BF16)
            narrow_kernel<1><<<blocks, 256, 0, g_pool.s_run>>>((const float *)atw(src[t], of), (uint16_t *)atw(dst[t], ob), n8);
          else
            narrow_kernel<0><<<blocks, 256, 0, g_pool.s_run>>>((const float *)atw(src[t], of), (uint16_t *)atw(dst[t], ob), n8);
        }
        FA_CUDA_CHECK(cudaGetLastError());
        count_launch(3);
      }
    }
    FA_CUDA_CHECK(cudaEventRecord(g_pool.run_done[g], g_pool.s_run));
    FA_CUDA_CHECK(cudaStreamWaitEvent(g_pool.s_out, g_pool.run_done[g], 0));
    FA_CUDA_CHECK(cudaMemcpyAsync(atw(O, ob), at(dOut, ob), nh * hb, cudaMemcpyDeviceToHost, g_pool.s_out));
    if (L) FA_CUDA_CHECK(cudaMemcpyAsync(atw(L, ol), at(dL, ol), nh * hl, cudaMemcpyDeviceToHost, g_pool.s_out));
    if (bwd && narrow) {
      FA_CUDA_CHECK(cudaMemcpyAsync(atw(dQ, ob), at(nQ, ob), nh * hb, cudaMemcpyDeviceToHost, g_pool.s_out));
      FA_CUDA_CHECK(cudaMemcpyAsync(atw(dK, ob), at(nK, ob), nh * hb, cudaMemcpyDeviceToHost, g_pool.s_out));
      FA_CUDA_CHECK(cudaMemcpyAsync(atw(dV, ob), at(nV, ob), nh * hb, cudaMemcpyDeviceToHost, g_pool.s_out));
    } else if (bwd) {
      FA_CUDA_CHECK(cudaMemcpyAsync(atw(dQ, of), at(gQ, of), nh * hf, cudaMemcpyDeviceToHost, g_pool.s_out));
      FA_CUDA_CHECK(cudaMemcpyAsync(atw(dK, of), at(gK, of), nh * hf, cudaMemcpyDeviceToHost, g_pool.s_out));
      FA_CUDA_CHECK(cudaMemcpyAsync(atw(dV, of), at(gV, of), nh * hf, cudaMemcpyDeviceToHost, g_pool.s_out));
    }
  }
  return FA_OK;
  };
  rc = enqueue_all();
  if (rc != FA_OK) return fail_after_enqueue(g_pool, rc);
  FA_CUDA_CHECK(cudaStreamSynchronize(g_pool.s_out));
  FA_CUDA_CHECK(cudaStreamSynchronize(g_pool.s_run));
  return FA_OK;
}

}  // namespace
}  // namespace fa

using namespace fa;

extern "C" {

int fa_host_attention_f32(int variant, const float *Q, const float *K, const float *V, float *O,
                          int N, int D, float scale, int is_causal) {
  FA_REQUIRE(variant >= 0 && variant <= 2, "variant must be 0, 1 or 2");
  FA_REQUIRE(Q && K && V && O && N >= 1 && D >= 1, "bad arguments");
  HostPool *pool_ptr = nullptr;
  int rc = current_pool(&pool_ptr);
  if (rc != FA_OK) return rc;
  HostPool &g_pool = *pool_ptr;
  std::lock_guard<std::mutex> lock(g_pool.mu);
  rc = g_pool.ensure_streams();
  if (rc != FA_OK) return rc;
  const size_t bytes = (size_t)N * D * sizeof(float);
  void *q, *k, *v, *o;
  if ((rc = g_pool.get(0, bytes, &q)) || (rc = g_pool.get(1, bytes, &k)) || (rc = g_pool.get(2, bytes, &v)) ||
      (rc = g_pool.get(3, bytes, &o)))
    return rc;
  cudaStream_t st = g_pool.s_run;
  auto enqueue_all = [&]() -> int {
    FA_CUDA_CHECK(cudaMemcpyAsync(q, Q, bytes, cudaMemcpyHostToDevice, st));
    FA_CUDA_CHECK(cudaMemcpyAsync(k, K, bytes, cudaMemcpyHostToDevice, st));
    FA_CUDA_CHECK(cudaMemcpyAsync(v, V, bytes, cudaMemcpyHostToDevice, st));
    const int lrc = launch_fp32(variant, (const float *)q, (const float *)k, (const float *)v, (float *)o, N, D, scale, 0, 0,
                                is_causal, 1, 1, st);
    if (lrc != FA_OK) return lrc;
    FA_CUDA_CHECK(cudaMemcpyAsync(O, o, bytes, cudaMemcpyDeviceToHost, st));
    return FA_OK;
  };
  rc = enqueue_all();
  if (rc != FA_OK) return fail_after_enqueue(g_pool, rc);
  FA_CUDA_CHECK(cudaStreamSynchronize(st));
  return FA_OK;
}

int fa_host_attention_half(const void *Q, const void *K, const void *V, void *O, float *L_out,
                           int N, int D, float scale, int is_causal, int B, int H, int dtype) {
  return host_half_impl(Q, K, V, nullptr, O, L_out, nullptr, nullptr, nullptr, N, D, scale, is_causal, B, H, dtype);
}

int fa_host_attention_fwd_bwd_half(const void *Q, const void *K, const void *V, const void *dO, void *O,
                                   float *L_out, float *dQ, float *dK, float *dV, int N, int D, float scale,
                                   int is_causal, int B, int H, int dtype) {
  FA_REQUIRE(dO != nullptr, "dO is null");
  return host_half_impl(Q, K, V, dO, O, L_out, dQ, dK, dV, N, D, scale, is_causal, B, H, dtype);
}

int fa_host_attention_fwd_bwd_half_ex(const void *Q, const void *K, const void *V, const void *dO, void *O,
                                      float *L_out, void *dQ, void *dK, void *dV, int N, int D, float scale,
                                      int is_causal, int B, int H, int dtype, int grad_dtype) {
  FA_REQUIRE(dO != nullptr, "dO is null");
  FA_REQUIRE(grad_dtype == -1 || grad_dtype == FA_DTYPE_FP16 || grad_dtype == FA_DTYPE_BF16,
             "grad_dtype must be -1 (fp32), FA_DTYPE_FP16 or FA_DTYPE_BF16");
  FA_REQUIRE(grad_dtype < 0 || ((size_t)N * D) % 8 == 0, "16-bit gradients need N * D to be a multiple of 8");
  return host_half_impl(Q, K, V, dO, O, L_out, dQ, dK, dV, N, D, scale, is_causal, B, H, dtype, grad_dtype);
}

void fa_host_release(void) {
  DeviceGuard guard;
  for (int dev = 0; dev < kMaxDevices; ++dev) {
    HostPool &pool = g_pools[dev];
    std::lock_guard<std::mutex> lock(pool.mu);
    if (!pool.s_in && !pool.ptr[0]) continue;
    if (cudaSetDevice(dev) != cudaSuccess) { (void)cudaGetLastError(); continue; }
    pool.release();
  }
}

}  // extern "C"
