// Development micro-benchmark (not part of the library): sustained rate of tcgen05.mma shapes
// used by the attention kernels, one CTA per SM, operands already in shared memory / TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flash_attention_metal_b200/csrc -o tools/mma_rate_probe tools/mma_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "sm100_ptx.cuh"
using namespace fa::ptx;

// mode 0: SS (A smem, B smem)   mode 1: TS (A tmem, B smem, MN-major B)
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) probe(long long *out, int iters) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  // zero the operand area so no NaNs slow anything down
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0;
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = a + 65536;
    constexpr uint32_t idesc = make_idesc(128, N, 1, 0, MODE == 1 ? 1 : 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (MODE == 0)
          mma_ss(tm + (it & 1) * 256, make_sdesc_sw128(a + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                 make_sdesc_sw128(b + (kk >> 2) * (N * 128) + (kk & 3) * 32, 16, 1024), idesc, kk > 0);
        else
          mma_ts(tm + (it & 1) * 256, tm + 128 + kk * 8, make_sdesc_sw128(b + kk * 2048, 16384, 1024), idesc, kk > 0);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int MODE, int N>
void run(const char *name, long long *d, int iters, int ctas) {
  cudaFuncSetAttribute(probe<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<MODE, N><<<ctas, 128, 200 * 1024>>>(d, iters);
  probe<MODE, N><<<ctas, 128, 200 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 8.0);
  printf("%-28s ctas=%3d  %.1f cycles per MMA (M=128,N=%d,K=16; ideal %d)  -> %.0f%% of peak  [%s]\n", name, ctas, per, N,
         N / 2, 100.0 * (N / 2) / per, cudaGetErrorString(e));
}

int main() {
  long long *d; cudaMalloc(&d, 64);
  for (int ctas : {1, 148}) {
    run<0, 128>("SS N=128 (Q K^T)", d, 2000, ctas);
    run<0, 64>("SS N=64 (bwd S^T, dP^T)", d, 2000, ctas);
    run<1, 128>("TS N=128 (P V, dV, dK, dQ)", d, 2000, ctas);
    run<1, 64>("TS N=64", d, 2000, ctas);
  }
  return 0;
}
