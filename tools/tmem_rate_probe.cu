// Development micro-benchmark: tcgen05.ld / tcgen05.st throughput (32x32b.x32) with 4 or 8 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flash_attention_metal_b200/csrc -o tools/tmem_rate_probe tools/tmem_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "sm100_ptx.cuh"
using namespace fa::ptx;

template <int MODE>  // 0 = ld, 1 = st, 2 = ld + 128 MUFU per 4 loads (softmax-like)
__global__ void __launch_bounds__(256, 1) probe(long long *out, int iters, int nwarps) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = i;
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (MODE == 1) { tmem_st32(base + c * 32, r); }
        else {
          tmem_ld32(base + c * 32, r);
          tmem_wait_ld();
          if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(ex2(__uint_as_float(r[i])));
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= r[i];
        }
      }
      if (MODE == 1) tmem_wait_st();
    }
    t1 = clock64();
  }
  if (acc == 0x12345678u) out[63] = acc;
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && warp < nwarps) out[warp] = t1 - t0;
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long *d; cudaMalloc(&d, 64 * 8);
  long long h[64];
  const int iters = 2000;
  const char *names[3] = {"tcgen05.ld 32x32b.x32 (4 KB per warp-instr)", "tcgen05.st 32x32b.x32", "ld + 32 ex2 per load"};
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 4, 8}) {
      cudaMemset(d, 0, 64 * 8);
      if (mode == 0) probe<0><<<148, 256>>>(d, iters, nw);
      if (mode == 1) probe<1><<<148, 256>>>(d, iters, nw);
      if (mode == 2) probe<2><<<148, 256>>>(d, iters, nw);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
      double per = (double)mx / (iters * 4.0);
      printf("%-46s warps/SM=%d: %.1f cycles per warp-instr, %.0f B/clk/SM  [%s]\n", names[mode], nw, per, nw * 4096.0 / per, cudaGetErrorString(e));
    }
  return 0;
}
