// Development micro-benchmark: issue cost (SM cycles per warp instruction) of the instructions the
// softmax loops are made of, alone and mixed with MUFU.EX2, at 1 and 2 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flash_attention_metal_b200/csrc -o tools/pipe_rate_probe tools/pipe_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "sm100_ptx.cuh"
using namespace fa::ptx;

enum { FFMA2, FADD2, FFMA, CVT, FMNMX, LEA, MUFU, MUFU_FFMA2, MUFU_FFMA, MUFU_CVT, MUFU_ALL, EMU_ONLY, NMODES };
const char *kNames[] = {"fma.f32x2", "add.f32x2", "fma.f32", "cvt.bf16x2", "max.f32", "shl+add", "ex2",
                        "ex2 + fma.f32x2 (1:1)", "ex2 + 2 fma.f32 (1:2)", "ex2 + cvt (1:1)",
                        "2 ex2 + fma2 + add2 + cvt (softmax pair)", "emulated pair only"};

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(long long *out, float *sink, int iters) {
  constexpr int NA = 16;
  uint64_t a[NA];
  float f[NA];
  uint32_t u[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    f[i] = -0.001f * (threadIdx.x + i);
    a[i] = pack_f32x2(f[i], f[i] * 0.5f);
    u[i] = threadIdx.x + i;
  }
  const uint64_t c2 = pack_f32x2(0.999f, 1.001f), d2 = pack_f32x2(-0.3f, -0.2f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (MODE == FFMA2) a[i] = fma_f32x2(a[i], c2, d2);
      if (MODE == FADD2) a[i] = add_f32x2(a[i], d2);
      if (MODE == FFMA) f[i] = fmaf(f[i], 0.999f, -0.3f);
      if (MODE == CVT) u[i] = pack2<1>(__uint_as_float(u[i]), f[i]);
      if (MODE == FMNMX) f[i] = fmaxf(f[i], __uint_as_float(u[(i + 1) % NA]));
      if (MODE == LEA) u[i] = (u[i] << 23) + u[(i + 1) % NA];
      if (MODE == MUFU) f[i] = ex2(f[i]);
      if (MODE == MUFU_FFMA2) { f[i] = ex2(f[i]); a[i] = fma_f32x2(a[i], c2, d2); }
      if (MODE == MUFU_FFMA) { f[i] = ex2(f[i]); u[i] = __float_as_uint(fmaf(__uint_as_float(u[i]), 0.999f, -0.3f));
                               a[i] = pack_f32x2(fmaf(lo_f32(a[i]), 0.999f, -0.3f), hi_f32(a[i])); }
      if (MODE == MUFU_CVT) { f[i] = ex2(f[i]); u[i] = pack2<1>(__uint_as_float(u[i]), lo_f32(a[i])); }
      if (MODE == MUFU_ALL) {
        const uint64_t x2 = fma_f32x2(a[i], c2, d2);
        const float p0 = ex2(lo_f32(x2)), p1 = ex2(hi_f32(x2));
        a[(i + 1) % NA] = add_f32x2(a[(i + 1) % NA], pack_f32x2(p0, p1));
        u[i] ^= pack2<1>(p0, p1);
      }
      if (MODE == EMU_ONLY) {
        const uint64_t x2 = fma_f32x2(a[i], c2, d2);
        const uint64_t e2 = exp2_emulated_x2(x2);
        a[(i + 1) % NA] = add_f32x2(a[(i + 1) % NA], e2);
        u[i] ^= pack2<1>(lo_f32(e2), hi_f32(e2));
      }
    }
  }
  long long t1 = clock64();
  float total = 0.f;
#pragma unroll
  for (int i = 0; i < NA; ++i) total += f[i] + lo_f32(a[i]) + hi_f32(a[i]) + __uint_as_float(u[i]);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = total;
}

template <int MODE>
void run(long long *d, float *sink) {
  const int iters = 4000;
  for (int nw : {4, 8, 16}) {
    probe<MODE><<<148, 32 * nw>>>(d, sink, iters);
    probe<MODE><<<148, 32 * nw>>>(d, sink, iters);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-42s %2d warps/SM: %6.2f cycles per unrolled step per warp (%.2f per scheduler-step)\n", kNames[MODE], nw,
           (double)h / iters / 16, (double)h / iters / 16 / (nw / 4));
  }
}

int main() {
  long long *d; float *sink; cudaMalloc(&d, 64); cudaMalloc(&sink, 148 * 512 * 4);
  run<FFMA2>(d, sink); run<FADD2>(d, sink); run<FFMA>(d, sink); run<CVT>(d, sink); run<FMNMX>(d, sink);
  run<LEA>(d, sink); run<MUFU>(d, sink); run<MUFU_FFMA2>(d, sink); run<MUFU_FFMA>(d, sink); run<MUFU_CVT>(d, sink);
  run<MUFU_ALL>(d, sink); run<EMU_ONLY>(d, sink);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
