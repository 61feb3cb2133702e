// flash_attn --config K [--gpus N]: the BASELINE.json configurations as presets of the harness.
//
//   1  the reference's CPU verifier alone, N=128 fp32 non-causal, faithful loop order (main.mm:128-159)
//   2  the harness sweep N=128..16384, fp32 + 16-bit, causal and not: four CSVs (main.mm:596-879)
//      -- handled in main.cpp, which loops over its own phases
//   3  flagship: causal bf16 forward + backward, B=1 H=16 N=16384 d=128 on one GPU
//   4  GPT-2-style B=8 H=12 N=4096 d=64 causal bf16: the 96 heads split over 1, 2, 4, 8 GPUs
//      (fa_mgpu_sharded_*: heads never interact, kernels.metal:622; call site main.mm:881-1204)
//   5  one long causal sequence N=131072..1M, d=128 bf16, ring / context-parallel attention over
//      the GPUs (fa_mgpu_ring_*), with a CPU spot check of sampled rows
// One process drives every GPU through the C ABI (fa_mgpu_*): no Python, no NCCL.  Each line prints
// ms, total and per-GPU TFLOP/s and the fraction of the bf16 tensor peak, measured (cuBLAS burst,
// MEASURED_PEAKS.json) and nominal (2250).  FLOPs: 4*B*H*N^2*d forward, halved when causal; backward
// 2.5x.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "flash_attn_b200.h"

namespace {

int g_bad = 0;

#define P_CUDA(x)                                                                     \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      std::fprintf(stderr, "CUDA Error: %s at %s\n", cudaGetErrorString(e_), #x);     \
      std::exit(1);                                                                   \
    }                                                                                 \
  } while (0)
#define P_FA(x)                                                                       \
  do {                                                                                \
    if ((x) != 0) {                                                                   \
      std::fprintf(stderr, "flash_attn Error in %s: %s\n", #x, fa_last_error());      \
      std::exit(1);                                                                   \
    }                                                                                 \
  } while (0)

uint16_t bf16_bits(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  x += 0x7fffu + ((x >> 16) & 1u);
  return (uint16_t)(x >> 16);
}
float bf16_value(uint16_t b) {
  uint32_t x = (uint32_t)b << 16;
  float f;
  std::memcpy(&f, &x, 4);
  return f;
}
// U(-1, 1) from a counter hash (splitmix64): any element of any tensor can be regenerated on the host
// for the spot checks without keeping host copies of gigabyte tensors
float uniform_at(uint64_t seed, uint64_t i) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + i + 0x632BE59BD9B4E019ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)((double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0);
}
uint16_t elem_bits(uint64_t seed, uint64_t i) { return bf16_bits(uniform_at(seed, i)); }

// device tensor [rows x D] whose row r is global row rows[r] of head h of the tensor with this seed
// (global element index = (h * n_total + row) * D + d)
void *upload_rows(uint64_t seed, int H, int64_t n_total, int D, const std::vector<int64_t> &rows) {
  const size_t n = rows.size();
  std::vector<uint16_t> host((size_t)H * n * D);
#pragma omp parallel for schedule(static)
  for (int64_t idx = 0; idx < (int64_t)((size_t)H * n); ++idx) {
    const int h = (int)(idx / (int64_t)n);
    const int64_t gr = rows[idx % n];
    for (int d = 0; d < D; ++d) host[(size_t)idx * D + d] = elem_bits(seed, ((uint64_t)h * n_total + gr) * D + d);
  }
  void *p = nullptr;
  P_CUDA(cudaMalloc(&p, host.size() * 2));
  P_CUDA(cudaMemcpy(p, host.data(), host.size() * 2, cudaMemcpyHostToDevice));
  return p;
}
std::vector<int64_t> iota_rows(int64_t first, int64_t count) {
  std::vector<int64_t> r((size_t)count);
  for (int64_t i = 0; i < count; ++i) r[(size_t)i] = first + i;
  return r;
}

double fwd_flops(double B, double H, double N, double D, bool causal) { return 4.0 * B * H * N * N * D * (causal ? 0.5 : 1.0); }

void print_header() {
  std::printf("%-34s %5s %9s %10s %12s %12s %9s %9s\n", "what", "GPUs", "N", "ms", "TFLOP/s", "per GPU", "%meas", "%nominal");
}
void print_row(const char *what, int gpus, int64_t N, double ms, double flops, double peak) {
  const double tf = flops / (ms * 1e-3) / 1e12;
  std::printf("%-34s %5d %9lld %10.3f %12.1f %12.1f %8.1f%% %8.1f%%\n", what, gpus, (long long)N, ms, tf, tf / gpus,
              100.0 * tf / gpus / peak, 100.0 * tf / gpus / 2250.0);
  std::fflush(stdout);
}

template <typename F>
double time_group_ms(fa_mgpu_t grp, F &&step, int warmup, int reps) {
  for (int i = 0; i < warmup; ++i) step();
  P_FA(fa_mgpu_synchronize(grp));
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) step();
  P_FA(fa_mgpu_synchronize(grp));
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
}

// fp64 attention of one query row against keys [0, n_keys) regenerated from the seeds
void cpu_row(uint64_t sq, uint64_t sk, uint64_t sv, int h, int64_t n_total, int D, int64_t row, int64_t n_keys, float scale,
             std::vector<double> &out) {
  std::vector<double> q(D), s((size_t)n_keys);
  for (int d = 0; d < D; ++d) q[d] = bf16_value(elem_bits(sq, ((uint64_t)h * n_total + row) * D + d));
  double mx = -INFINITY;
#pragma omp parallel for reduction(max : mx) schedule(static)
  for (int64_t j = 0; j < n_keys; ++j) {
    double acc = 0;
    for (int d = 0; d < D; ++d) acc += q[d] * bf16_value(elem_bits(sk, ((uint64_t)h * n_total + j) * D + d));
    s[(size_t)j] = acc * scale;
    mx = std::max(mx, s[(size_t)j]);
  }
  out.assign(D, 0.0);
  double den = 0;
  for (int64_t j = 0; j < n_keys; ++j) {
    const double p = std::exp(s[(size_t)j] - mx);
    den += p;
    for (int d = 0; d < D; ++d) out[d] += p * bf16_value(elem_bits(sv, ((uint64_t)h * n_total + j) * D + d));
  }
  for (int d = 0; d < D; ++d) out[d] /= den;
}

void config3(double peak, int reps) {
  const int B = 1, H = 16, N = 16384, D = 128;
  const float scale = 1.0f / std::sqrt((float)D);
  std::printf("\n== config 3: causal bf16 forward + backward, B=%d H=%d N=%d d=%d, one GPU ==\n", B, H, N, D);
  P_CUDA(cudaSetDevice(0));
  const auto rows = iota_rows(0, N);
  void *Q = upload_rows(42, H, N, D, rows), *K = upload_rows(43, H, N, D, rows), *V = upload_rows(44, H, N, D, rows),
       *dO = upload_rows(45, H, N, D, rows);
  const size_t e = (size_t)B * H * N * D;
  void *O;
  float *L, *dQ, *dK, *dV;
  void *ws;
  const size_t wsb = fa_workspace_bytes_backward(N, D, B, H);
  P_CUDA(cudaMalloc(&O, e * 2)); P_CUDA(cudaMalloc(&L, (size_t)B * H * N * 4));
  P_CUDA(cudaMalloc(&dQ, e * 4)); P_CUDA(cudaMalloc(&dK, e * 4)); P_CUDA(cudaMalloc(&dV, e * 4)); P_CUDA(cudaMalloc(&ws, wsb));
  const int64_t hs = (int64_t)N * D, bs = hs * H;
  auto fwd = [&] { P_FA(flash_attention_v4_half(Q, K, V, O, N, D, scale, bs, hs, L, 1, B, H, FA_DTYPE_BF16, nullptr)); };
  auto bwd = [&] { P_FA(flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, N, D, scale, bs, hs, 1, B, H, FA_DTYPE_BF16, ws, wsb, nullptr)); };
  auto timed = [&](auto &&f) {
    cudaEvent_t a, b;
    P_CUDA(cudaEventCreate(&a)); P_CUDA(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    P_CUDA(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    P_CUDA(cudaEventRecord(b));
    P_CUDA(cudaEventSynchronize(b));
    float ms;
    P_CUDA(cudaEventElapsedTime(&ms, a, b));
    return (double)ms / reps;
  };
  print_header();
  const double ff = fwd_flops(B, H, N, D, true);
  print_row("forward (flash_attention_v4_half)", 1, N, timed(fwd), ff, peak);
  print_row("backward (3 kernels)", 1, N, timed(bwd), 2.5 * ff, peak);
  print_row("forward + backward", 1, N, timed([&] { fwd(); bwd(); }), 3.5 * ff, peak);
  // spot check: three rows of head H-1 against the fp64 row reference (causal: keys <= row)
  std::vector<uint16_t> orow(D);
  double worst = 0;
  for (int64_t r : {(int64_t)0, (int64_t)N / 2 + 77, (int64_t)N - 1}) {
    std::vector<double> want;
    cpu_row(42, 43, 44, H - 1, N, D, r, r + 1, scale, want);
    P_CUDA(cudaMemcpy(orow.data(), (char *)O + (((size_t)(H - 1) * N + r) * D) * 2, D * 2, cudaMemcpyDeviceToHost));
    for (int d = 0; d < D; ++d) worst = std::max(worst, std::abs(bf16_value(orow[d]) - want[d]));
  }
  std::printf("spot check of 3 rows vs fp64 CPU rows: max-abs %.3e (<= 2e-2) %s\n", worst, worst <= 2e-2 ? "PASSED" : "FAILED");
  if (!(worst <= 2e-2)) ++g_bad;
  for (void *p : {Q, K, V, dO, O, (void *)L, (void *)dQ, (void *)dK, (void *)dV, ws}) cudaFree(p);
}

void config4(int max_gpus, double peak, int reps) {
  const int B = 8, Hh = 12, N = 4096, D = 64, heads = B * Hh;
  const float scale = 1.0f / std::sqrt((float)D);
  std::printf("\n== config 4: GPT-2-style B=%d H=%d N=%d d=%d causal bf16 forward + backward, %d heads split over the GPUs ==\n", B,
              Hh, N, D, heads);
  print_header();
  const double flops = 3.5 * fwd_flops(B, Hh, N, D, true);
  double one_gpu_ms = 0;
  for (int g = 1; g <= max_gpus; g *= 2) {
    std::vector<int> devs(g), hpd(g);
    for (int i = 0; i < g; ++i) { devs[i] = i; hpd[i] = heads / g + (i < heads % g ? 1 : 0); }
    fa_mgpu_t grp;
    P_FA(fa_mgpu_create(&grp, devs.data(), g));
    std::vector<void *> Q(g), K(g), V(g), dO(g), O(g);
    std::vector<float *> L(g), dQ(g), dK(g), dV(g);
    int h0 = 0;
    for (int i = 0; i < g; ++i) {
      P_CUDA(cudaSetDevice(i));
      // head h of the global problem is "head h" of one [heads, N, D] tensor
      std::vector<int64_t> rows;
      rows.reserve((size_t)hpd[i] * N);
      const auto all = iota_rows(0, N);
      auto up = [&](uint64_t seed) {
        // upload_rows generates [H, rows]; shift the head index by h0 through the seed-independent offset
        std::vector<uint16_t> host((size_t)hpd[i] * N * D);
#pragma omp parallel for schedule(static)
        for (int64_t idx = 0; idx < (int64_t)hpd[i] * N; ++idx)
          for (int d = 0; d < D; ++d) host[(size_t)idx * D + d] = elem_bits(seed, ((uint64_t)h0 * N + idx) * D + d);
        void *p;
        P_CUDA(cudaMalloc(&p, host.size() * 2));
        P_CUDA(cudaMemcpy(p, host.data(), host.size() * 2, cudaMemcpyHostToDevice));
        return p;
      };
      Q[i] = up(52); K[i] = up(53); V[i] = up(54); dO[i] = up(55);
      const size_t e = (size_t)hpd[i] * N * D;
      P_CUDA(cudaMalloc(&O[i], e * 2)); P_CUDA(cudaMalloc(&L[i], (size_t)hpd[i] * N * 4));
      P_CUDA(cudaMalloc(&dQ[i], e * 4)); P_CUDA(cudaMalloc(&dK[i], e * 4)); P_CUDA(cudaMalloc(&dV[i], e * 4));
      h0 += hpd[i];
    }
    auto step = [&] {
      P_FA(fa_mgpu_sharded_forward(grp, Q.data(), K.data(), V.data(), O.data(), L.data(), N, D, scale, 1, hpd.data(), FA_DTYPE_BF16));
      P_FA(fa_mgpu_sharded_backward(grp, Q.data(), K.data(), V.data(), O.data(), dO.data(), L.data(), dQ.data(), dK.data(),
                                    dV.data(), N, D, scale, 1, hpd.data(), FA_DTYPE_BF16));
    };
    const double ms = time_group_ms(grp, step, 3, reps);
    if (g == 1) one_gpu_ms = ms;
    char what[64];
    std::snprintf(what, sizeof(what), "sharded fwd+bwd (%.2fx of 1 GPU)", one_gpu_ms / ms);
    print_row(what, g, N, ms, flops, peak);
    // spot check on the last device: last row of its last head (sees every key)
    {
      const int i = g - 1, hl = hpd[i] - 1, hglob = heads - 1;
      std::vector<double> want;
      cpu_row(52, 53, 54, hglob, N, D, N - 1, N, scale, want);
      std::vector<uint16_t> orow(D);
      P_CUDA(cudaSetDevice(i));
      P_CUDA(cudaMemcpy(orow.data(), (char *)O[i] + (((size_t)hl * N + N - 1) * D) * 2, D * 2, cudaMemcpyDeviceToHost));
      double worst = 0;
      for (int d = 0; d < D; ++d) worst = std::max(worst, std::abs(bf16_value(orow[d]) - want[d]));
      if (!(worst <= 2e-2)) { std::printf("  spot check FAILED on %d GPUs: max-abs %.3e\n", g, worst); ++g_bad; }
    }
    for (int i = 0; i < g; ++i) {
      P_CUDA(cudaSetDevice(i));
      for (void *p : {Q[i], K[i], V[i], dO[i], O[i], (void *)L[i], (void *)dQ[i], (void *)dK[i], (void *)dV[i]}) cudaFree(p);
    }
    P_FA(fa_mgpu_destroy(grp));
  }
}

void config5(int gpus, double peak, int reps, int64_t max_n) {
  const int D = 128;
  const float scale = 1.0f / std::sqrt((float)D);
  std::printf("\n== config 5: one causal bf16 sequence, d=%d, ring / context-parallel attention over %d GPUs (zig-zag chunks) ==\n", D, gpus);
  print_header();
  std::vector<int> devs(gpus);
  for (int i = 0; i < gpus; ++i) devs[i] = i;
  fa_mgpu_t grp;
  P_FA(fa_mgpu_create(&grp, devs.data(), gpus));
  for (int64_t N : {(int64_t)131072, (int64_t)262144, (int64_t)524288, (int64_t)1048576}) {
    if (N > max_n) break;
    const int H = N <= 131072 ? 4 : (N <= 262144 ? 2 : 1);
    const int n_local = (int)(N / gpus);
    std::vector<void *> Q(gpus), K(gpus), V(gpus), dO(gpus), O(gpus);
    std::vector<float *> L(gpus), dQ(gpus), dK(gpus), dV(gpus);
    std::vector<std::vector<int64_t>> rows(gpus);
    for (int i = 0; i < gpus; ++i) {
      int64_t first[2];
      int cnt[2];
      P_FA(fa_ring_local_rows(i, gpus, n_local, 1, first, cnt));
      for (int c = 0; c < 2; ++c)
        for (int r = 0; r < cnt[c]; ++r) rows[i].push_back(first[c] + r);
      P_CUDA(cudaSetDevice(i));
      Q[i] = upload_rows(62, H, N, D, rows[i]); K[i] = upload_rows(63, H, N, D, rows[i]);
      V[i] = upload_rows(64, H, N, D, rows[i]); dO[i] = upload_rows(65, H, N, D, rows[i]);
      const size_t e = (size_t)H * n_local * D;
      P_CUDA(cudaMalloc(&O[i], e * 2)); P_CUDA(cudaMalloc(&L[i], (size_t)H * n_local * 4));
      P_CUDA(cudaMalloc(&dQ[i], e * 4)); P_CUDA(cudaMalloc(&dK[i], e * 4)); P_CUDA(cudaMalloc(&dV[i], e * 4));
    }
    auto fwd = [&] { P_FA(fa_mgpu_ring_forward(grp, Q.data(), K.data(), V.data(), O.data(), L.data(), n_local, D, H, scale, 1, FA_DTYPE_BF16)); };
    auto bwd = [&] {
      P_FA(fa_mgpu_ring_backward(grp, Q.data(), K.data(), V.data(), O.data(), dO.data(), L.data(), dQ.data(), dK.data(), dV.data(),
                                 n_local, D, H, scale, 1, FA_DTYPE_BF16));
    };
    const double ff = fwd_flops(1, H, (double)N, D, true);
    const int r = N >= 524288 ? std::max(1, reps / 4) : reps;
    char what[64];
    std::snprintf(what, sizeof(what), "ring forward, H=%d", H);
    print_row(what, gpus, N, time_group_ms(grp, fwd, 1, r), ff, peak);
    std::snprintf(what, sizeof(what), "ring forward + backward, H=%d", H);
    print_row(what, gpus, N, time_group_ms(grp, [&] { fwd(); bwd(); }, 1, r), 3.5 * ff, peak);
    // spot check: on the last GPU, its first local row (global chunk P-1) and its last one (global chunk P)
    double worst = 0;
    {
      const int i = gpus - 1;
      P_CUDA(cudaSetDevice(i));
      for (int lr : {0, n_local - 1}) {
        const int64_t gr = rows[i][(size_t)lr];
        std::vector<double> want;
        cpu_row(62, 63, 64, H - 1, N, D, gr, gr + 1, scale, want);
        std::vector<uint16_t> orow(D);
        P_CUDA(cudaMemcpy(orow.data(), (char *)O[i] + (((size_t)(H - 1) * n_local + lr) * D) * 2, D * 2, cudaMemcpyDeviceToHost));
        for (int d = 0; d < D; ++d) worst = std::max(worst, std::abs(bf16_value(orow[d]) - want[d]));
      }
    }
    std::printf("  spot check of 2 rows vs fp64 CPU rows: max-abs %.3e (<= 2e-2) %s\n", worst, worst <= 2e-2 ? "PASSED" : "FAILED");
    if (!(worst <= 2e-2)) ++g_bad;
    for (int i = 0; i < gpus; ++i) {
      P_CUDA(cudaSetDevice(i));
      for (void *p : {Q[i], K[i], V[i], dO[i], O[i], (void *)L[i], (void *)dQ[i], (void *)dK[i], (void *)dV[i]}) cudaFree(p);
    }
  }
  P_FA(fa_mgpu_destroy(grp));
}

}  // namespace

// Entry point used by main.cpp.  Returns the number of failed spot checks.
int run_presets(int config, int gpus, double peak_tflops, int reps, long long max_n) {
  const int visible = fa_device_count();
  if (visible < 1) {
    std::fprintf(stderr, "Error: No CUDA device found.\n");
    return 1;
  }
  if (gpus < 1) gpus = 1;
  if (gpus > visible) {
    std::fprintf(stderr, "Error: --gpus %d but only %d CUDA devices are visible.\n", gpus, visible);
    return 1;
  }
  std::printf("flash_attn presets: config %d on %d of %d visible GPUs; tensor peak %.1f TFLOP/s measured (cuBLAS bf16 burst), 2250 nominal\n",
              config, gpus, visible, peak_tflops);
  if (config == 3) config3(peak_tflops, reps);
  else if (config == 4) config4(gpus, peak_tflops, reps);
  else if (config == 5) config5(gpus, peak_tflops, reps, max_n);
  std::printf("%s\n", g_bad ? "PRESET FAILED" : "PRESET PASSED");
  return g_bad;
}
