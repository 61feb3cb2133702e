// Development micro-benchmark: cycles of the forward softmax inner loop (128 columns per thread, one
// warp per scheduler) with 0, 1/4, 1/3, 1/2 of the exponentials emulated on the FMA pipe.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "sm100_ptx.cuh"
using namespace fa::ptx;

template <int EMU, int NW>
__global__ void __launch_bounds__(32 * NW, 1) probe(long long *out, float *sink, int iters, float scale, float negm) {
  uint32_t s[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i) s[c][i] = __float_as_uint(-0.01f * (threadIdx.x + c * 32 + i));
  const uint64_t scale2 = pack_f32x2(scale, scale), negm2 = pack_f32x2(negm, negm);
  float total = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint64_t sum2[2] = {0ull, 0ull};
    uint32_t pk[64];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const uint64_t x2 = fma_f32x2(pack_u32x2(s[c][i], s[c][i + 1]), scale2, negm2);
        float p0, p1;
        if (EMU > 0 && ((c * 16 + (i >> 1)) % (EMU > 0 ? EMU : 1)) == (EMU > 0 ? EMU : 1) - 1) {
          const uint64_t e = exp2_emulated_x2(x2);
          p0 = lo_f32(e); p1 = hi_f32(e);
        } else {
          p0 = ex2(lo_f32(x2)); p1 = ex2(hi_f32(x2));
        }
        sum2[(i >> 1) & 1] = add_f32x2(sum2[(i >> 1) & 1], pack_f32x2(p0, p1));
        pk[c * 16 + (i >> 1)] = pack2<1>(p0, p1);
      }
    const uint64_t st2 = add_f32x2(sum2[0], sum2[1]);
    total += lo_f32(st2) + hi_f32(st2);
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) x ^= pk[i];
    // feed the result back so iterations cannot be merged
#pragma unroll
    for (int i = 0; i < 32; ++i) s[it & 3][i] ^= (x & 1u);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = total;
}

template <int EMU, int NW>
void run(long long *d, float *sink) {
  const int iters = 2000;
  probe<EMU, NW><<<148, 32 * NW>>>(d, sink, iters, 0.1275f, -0.3f);
  probe<EMU, NW><<<148, 32 * NW>>>(d, sink, iters, 0.1275f, -0.3f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("EMU=%d warps/SM=%d (%d per scheduler): %.0f cycles per 128-column row pass  [%s]\n", EMU, NW, NW / 4, (double)h / iters,
         cudaGetErrorString(e));
}

// pure MUFU.EX2 throughput: 128 independent exponentials per pass
template <int NW>
__global__ void __launch_bounds__(32 * NW, 1) mufu_probe(long long *out, float *sink, int iters) {
  float v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = -0.01f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 128; ++i) v[i] = ex2(v[i]) - 1.5f;
  }
  long long t1 = clock64();
  float total = 0.f;
#pragma unroll
  for (int i = 0; i < 128; ++i) total += v[i];
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = total;
}
template <int NW>
void run_mufu(long long *d, float *sink) {
  const int iters = 2000;
  mufu_probe<NW><<<148, 32 * NW>>>(d, sink, iters);
  mufu_probe<NW><<<148, 32 * NW>>>(d, sink, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("MUFU.EX2 + FADD only, warps/SM=%d: %.0f cycles per 128 exponentials per thread  [%s]\n", NW, (double)h / iters, cudaGetErrorString(e));
}

int main() {
  long long *d; float *sink; cudaMalloc(&d, 64); cudaMalloc(&sink, 148 * 256 * 4);
  run_mufu<4>(d, sink); run_mufu<8>(d, sink); run_mufu<16>(d, sink);
  run<0, 4>(d, sink); run<4, 4>(d, sink); run<3, 4>(d, sink); run<2, 4>(d, sink);
  run<0, 8>(d, sink); run<4, 8>(d, sink); run<3, 8>(d, sink); run<2, 8>(d, sink);
  return 0;
}
