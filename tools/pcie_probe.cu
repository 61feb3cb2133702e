// pcie_probe: copy-only ceiling of the host-buffer (end-to-end) path -- pinned host -> device and device -> host
// at the same time on two streams, plain cudaMemcpyAsync, no kernels.  Run one instance per GPU concurrently
// (e.g. `for i in 0 1 2 3 4 5 6 7; do ./pcie_probe $i & done; wait`) to see what the HOST sustains when every
// rank copies at once: that is the bound on `e2e` in bench.py (which measures the same thing in-process as
// e2e.copy_ceiling_gbs).  Build: nvcc -O2 -o tools/pcie_probe tools/pcie_probe.cu
// usage: pcie_probe [device] [h2d_MB] [d2h_MB] [reps]     (defaults: 0 256 449 10 = the flagship step's bytes)
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      std::fprintf(stderr, "CUDA error %s at %s\n", cudaGetErrorString(e_), #x);           \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

int main(int argc, char **argv) {
  const int dev = argc > 1 ? std::atoi(argv[1]) : 0;
  const size_t in_bytes = (size_t)(argc > 2 ? std::atoi(argv[2]) : 256) << 20;
  const size_t out_bytes = (size_t)(argc > 3 ? std::atoi(argv[3]) : 449) << 20;
  const int reps = argc > 4 ? std::atoi(argv[4]) : 10;
  CK(cudaSetDevice(dev));
  void *h_in, *h_out, *d_in, *d_out;
  CK(cudaMallocHost(&h_in, in_bytes));
  CK(cudaMallocHost(&h_out, out_bytes));
  CK(cudaMalloc(&d_in, in_bytes));
  CK(cudaMalloc(&d_out, out_bytes));
  cudaStream_t s_in, s_out;
  CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  auto once = [&]() -> int {
    CK(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, s_in));
    CK(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s_out));
    return 0;
  };
  if (once()) return 1;
  CK(cudaDeviceSynchronize());
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i)
    if (once()) return 1;
  CK(cudaDeviceSynchronize());
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
  std::printf("device %d: %.1f MB in + %.1f MB out per step, %.3f ms -> %.1f GB/s both directions together (h2d %.1f, d2h %.1f)\n", dev,
              in_bytes / 1e6, out_bytes / 1e6, s * 1e3, (in_bytes + out_bytes) / s / 1e9, in_bytes / s / 1e9, out_bytes / s / 1e9);
  return 0;
}
