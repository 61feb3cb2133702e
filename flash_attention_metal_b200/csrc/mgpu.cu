// fa_mgpu_*: one process driving several GPUs of one box (last table row; the
// reference's harness is a single process, main.mm:881-1204).  The group owns one stream per device,
// grow-only scratch, and -- for ring attention -- one Ring per device wired to its peers with plain
// peer access (no NCCL, no IPC: one address space).  Every call only ENQUEUES work on the group's
// streams -- every device's share from its own worker thread, concurrently; fa_mgpu_synchronize waits for it.
//
//   * B x H sharding (BASELINE config 4): heads never interact (kernels.metal:622), so device i runs
//     its own heads with the ordinary kernels and nothing is exchanged.
//   * ring / context-parallel attention (config 5): the same driver as fa_ring_attention_* with the
//     PEER transport -- copy-engine pulls over NVLink, flags driven by stream memory operations.
#include <cuda_runtime.h>

#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "fa_internal.h"
#include "ring.h"

namespace fa {

// One host thread per device, alive for the life of the group.  A group call hands every rank's share of the
// enqueue to its worker and waits for all of them: the ranks enqueue CONCURRENTLY, like the processes of the
// one-process-per-GPU model.  That is not an optimisation detail.  A ring call puts dozens of flag waits and
// copies into a rank's streams; enqueued one rank after the other from a single thread, rank 0's streams sit
// parked on flags until rank P-1 has been enqueued, the call's latency grows with P x (host time per rank)
// (measured on 8 GPUs: 5.9 ms for a 1.9 ms ring forward), and a few calls in a row overflow a parked stream's
// hardware queue, at which point the enqueueing thread itself blocks -- and nobody is left to enqueue the rank
// that would write the flag (a real hang on 8 GPUs).  With one thread per rank a blocked enqueue only blocks
// its own rank.
struct Worker {
  std::thread thread;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = false, quit = false;
  int rc = FA_OK;
  long launches = 0;
  char err[512] = "";
};

struct MgpuGroup {
  int n = 0;
  int devices[kRingMaxWorld] = {};
  Ring *rings[kRingMaxWorld] = {};
  cudaStream_t streams[kRingMaxWorld] = {};
  void *ws[kRingMaxWorld] = {};      // per-device scratch (ring workspace / backward delta)
  size_t ws_cap[kRingMaxWorld] = {};
  Worker workers[kRingMaxWorld];
};

namespace {

void worker_main(Worker *w, int device) {
  cudaSetDevice(device);
  for (;;) {
    std::function<int()> job;
    {
      std::unique_lock<std::mutex> lock(w->mu);
      w->cv.wait(lock, [w] { return w->has_job || w->quit; });
      if (w->quit) return;
      job = std::move(w->job);
      w->has_job = false;
    }
    fa_reset_launch_count();
    const int rc = job();
    {
      std::lock_guard<std::mutex> lock(w->mu);
      w->rc = rc;
      w->launches = fa_launch_count();
      if (rc != FA_OK) strncpy(w->err, fa_last_error(), sizeof(w->err) - 1);
      w->done = true;
    }
    w->cv.notify_all();
  }
}

// run job(i) on worker i for every device of the group, concurrently; first error wins
int run_on_all(MgpuGroup *g, const std::function<int(int)> &job) {
  for (int i = 0; i < g->n; ++i) {
    Worker &w = g->workers[i];
    {
      std::lock_guard<std::mutex> lock(w.mu);
      w.job = [&job, i] { return job(i); };
      w.has_job = true;
      w.done = false;
    }
    w.cv.notify_all();
  }
  int rc = FA_OK;
  for (int i = 0; i < g->n; ++i) {
    Worker &w = g->workers[i];
    std::unique_lock<std::mutex> lock(w.mu);
    w.cv.wait(lock, [&w] { return w.done; });
    count_launch((int)w.launches);
    if (w.rc != FA_OK && rc == FA_OK) rc = set_error(w.rc, "device %d: %s", g->devices[i], w.err);
  }
  return rc;
}

void stop_workers(MgpuGroup *g) {
  for (int i = 0; i < g->n; ++i) {
    Worker &w = g->workers[i];
    if (!w.thread.joinable()) continue;
    {
      std::lock_guard<std::mutex> lock(w.mu);
      w.quit = true;
    }
    w.cv.notify_all();
    w.thread.join();
  }
}

}  // namespace

// Grow every rank's peer-visible window (called from the first rank's ring call that needs more).
int mgpu_grow_windows(MgpuGroup *g, size_t bytes) {
  const size_t cap = (bytes + (size_t(1) << 21) - 1) & ~((size_t(1) << 21) - 1);
  for (int i = 0; i < g->n; ++i) {  // nothing may still be reading the old windows
    FA_CUDA_CHECK(cudaSetDevice(g->devices[i]));
    FA_CUDA_CHECK(cudaDeviceSynchronize());
  }
  for (int i = 0; i < g->n; ++i) {
    Ring *r = g->rings[i];
    FA_CUDA_CHECK(cudaSetDevice(g->devices[i]));
    if (r->data) cudaFree(r->data);
    r->data = nullptr;
    r->data_cap = 0;
    void *p = nullptr;
    FA_CUDA_CHECK(cudaMalloc(&p, cap));
    r->data = reinterpret_cast<char *>(p);
    r->data_cap = cap;
  }
  for (int i = 0; i < g->n; ++i)
    for (int j = 0; j < g->n; ++j) g->rings[i]->peer_data[j] = g->rings[j]->data;
  return FA_OK;
}

namespace {

// Make every device's scratch at least `bytes` and every rank's peer window at least `window_bytes`
// BEFORE anything of the call is enqueued.  Growing means freeing, and cudaFree / cudaDeviceSynchronize
// wait for the whole device: once one rank's ring work is in its streams (blocked on flags that a later
// rank has yet to write) such a wait would never return.  At the start of a call every earlier call has
// been enqueued for all ranks, so the wait is safe there -- and only there.
int group_reserve(MgpuGroup *g, size_t bytes, size_t window_bytes) {
  bool grow = false;
  for (int i = 0; i < g->n; ++i) grow = grow || g->ws_cap[i] < bytes;
  if (grow) {
    for (int i = 0; i < g->n; ++i) {
      FA_CUDA_CHECK(cudaSetDevice(g->devices[i]));
      FA_CUDA_CHECK(cudaDeviceSynchronize());
    }
    for (int i = 0; i < g->n; ++i) {
      if (g->ws_cap[i] >= bytes) continue;
      FA_CUDA_CHECK(cudaSetDevice(g->devices[i]));
      if (g->ws[i]) cudaFree(g->ws[i]);
      g->ws[i] = nullptr;
      g->ws_cap[i] = 0;
      FA_CUDA_CHECK(cudaMalloc(&g->ws[i], bytes));
      g->ws_cap[i] = bytes;
    }
  }
  if (window_bytes > 0 && g->rings[0]->data_cap < window_bytes) return mgpu_grow_windows(g, window_bytes);
  return FA_OK;
}

void group_free(MgpuGroup *g) {
  if (!g) return;
  stop_workers(g);
  for (int i = 0; i < g->n; ++i) {
    cudaSetDevice(g->devices[i]);
    cudaDeviceSynchronize();
  }
  for (int i = 0; i < g->n; ++i) {
    cudaSetDevice(g->devices[i]);
    if (g->rings[i]) {
      g->rings[i]->group = nullptr;  // ring_free then frees the windows like a stand-alone ring's
      for (int p = 0; p < kRingMaxWorld; ++p) { g->rings[i]->peer_data[p] = nullptr; g->rings[i]->peer_flags[p] = nullptr; }
      ring_free(g->rings[i]);
    }
    if (g->ws[i]) cudaFree(g->ws[i]);
    if (g->streams[i]) cudaStreamDestroy(g->streams[i]);
  }
  (void)cudaGetLastError();
  delete g;
}

}  // namespace
}  // namespace fa

using namespace fa;

extern "C" {

int fa_mgpu_create(void **out, const int *devices, int n_devices) {
  DeviceGuard guard;
  FA_REQUIRE(out && devices && n_devices >= 1 && n_devices <= kRingMaxWorld, "bad device list (1..%d devices)", kRingMaxWorld);
  int visible = 0;
  FA_CUDA_CHECK(cudaGetDeviceCount(&visible));
  for (int i = 0; i < n_devices; ++i) {
    FA_REQUIRE(devices[i] >= 0 && devices[i] < visible, "device %d is not visible (%d devices)", devices[i], visible);
    for (int j = 0; j < i; ++j) FA_REQUIRE(devices[i] != devices[j], "device %d is listed twice", devices[i]);
  }
  MgpuGroup *g = new MgpuGroup();
  g->n = n_devices;
  auto fail = [&](int rc) { group_free(g); return rc; };
  for (int i = 0; i < n_devices; ++i) {
    g->devices[i] = devices[i];
    if (cudaSetDevice(devices[i]) != cudaSuccess) return fail(set_error(FA_ERR_CUDA, "cudaSetDevice(%d) failed", devices[i]));
    for (int j = 0; j < n_devices; ++j) {
      if (j == i) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
      if (!can) return fail(set_error(FA_ERR_UNSUPPORTED, "device %d cannot access device %d as a peer", devices[i], devices[j]));
      const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(set_error(FA_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", devices[i], devices[j], cudaGetErrorString(e)));
      (void)cudaGetLastError();
    }
    if (cudaStreamCreateWithFlags(&g->streams[i], cudaStreamNonBlocking) != cudaSuccess)
      return fail(set_error(FA_ERR_CUDA, "cudaStreamCreate failed on device %d", devices[i]));
    Ring *r = new Ring();
    g->rings[i] = r;
    r->group = g;
    r->rank = i; r->world = n_devices; r->device = devices[i];
    r->transport = FA_RING_TRANSPORT_PEER;
    int rc = ring_init_streams(r);
    if (rc == FA_OK) rc = ring_alloc_flags(r);
    // nothing may be loaded lazily later: a load synchronises the context, and the context may then hold a
    // stream parked on a flag that a later rank's enqueue has yet to write
    if (rc == FA_OK) rc = preload_kernels();
    if (rc != FA_OK) return fail(rc);
  }
  for (int i = 0; i < n_devices; ++i)
    for (int j = 0; j < n_devices; ++j) g->rings[i]->peer_flags[j] = g->rings[j]->flags;
  for (int i = 0; i < n_devices; ++i) g->workers[i].thread = std::thread(worker_main, &g->workers[i], devices[i]);
  *out = g;
  return FA_OK;
}

// Development aid (not in the public header): the ring object of the index-th device (for fa_debug_ring_flags)
void *fa_debug_mgpu_ring(void *group, int index) {
  MgpuGroup *g = reinterpret_cast<MgpuGroup *>(group);
  return (g && index >= 0 && index < g->n) ? g->rings[index] : nullptr;
}

int fa_mgpu_destroy(void *group) {
  DeviceGuard guard;
  group_free(reinterpret_cast<MgpuGroup *>(group));
  return FA_OK;
}

int fa_mgpu_device_count(void *group) {
  MgpuGroup *g = reinterpret_cast<MgpuGroup *>(group);
  return g ? g->n : 0;
}

fa_stream_t fa_mgpu_stream(void *group, int index) {
  MgpuGroup *g = reinterpret_cast<MgpuGroup *>(group);
  return (g && index >= 0 && index < g->n) ? (fa_stream_t)g->streams[index] : nullptr;
}

int fa_mgpu_synchronize(void *group) {
  DeviceGuard guard;
  MgpuGroup *g = reinterpret_cast<MgpuGroup *>(group);
  FA_REQUIRE(g, "null group");
  for (int i = 0; i < g->n; ++i) {
    FA_CUDA_CHECK(cudaSetDevice(g->devices[i]));
    FA_CUDA_CHECK(cudaStreamSynchronize(g->streams[i]));
  }
  return FA_OK;
}

int fa_mgpu_sharded_forward(void *group, const void *const *Q, const void *const *K, const void *const *V,
                            void *const *O, float *const *L, int N, int D, float scale, int is_causal,
                            const int *heads, int dtype) {
  DeviceGuard guard;
  MgpuGroup *g = reinterpret_cast<MgpuGroup *>(group);
  FA_REQUIRE(g && Q && K && V && O && heads, "null argument");
  return run_on_all(g, [&](int i) -> int {
    if (heads[i] <= 0) return FA_OK;
    const int64_t hs = (int64_t)N * D;
    return launch_fwd_tc(Q[i], K[i], V[i], O[i], L ? L[i] : nullptr, N, D, scale, hs * heads[i], hs, is_causal, 1, heads[i],
                         dtype, g->streams[i]);
  });
}

int fa_mgpu_sharded_backward(void *group, const void *const *Q, const void *const *K, const void *const *V,
                             const void *const *O, const void *const *dO, const float *const *L, float *const *dQ,
                             float *const *dK, float *const *dV, int N, int D, float scale, int is_causal,
                             const int *heads, int dtype) {
  DeviceGuard guard;
  MgpuGroup *g = reinterpret_cast<MgpuGroup *>(group);
  FA_REQUIRE(g && Q && K && V && O && dO && L && dQ && dK && dV && heads, "null argument");
  size_t need = 0;
  for (int i = 0; i < g->n; ++i)
    if (heads[i] > 0) need = std::max(need, fa_workspace_bytes_backward(N, D, 1, heads[i]));
  int rc = group_reserve(g, need, 0);
  if (rc != FA_OK) return rc;
  return run_on_all(g, [&](int i) -> int {
    if (heads[i] <= 0) return FA_OK;
    const int64_t hs = (int64_t)N * D;
    const size_t wsb = fa_workspace_bytes_backward(N, D, 1, heads[i]);
    return launch_bwd_tc(Q[i], K[i], V[i], O[i], dO[i], L[i], dQ[i], dK[i], dV[i], N, D, scale, hs * heads[i], hs, is_causal, 1,
                         heads[i], dtype, g->ws[i], wsb, g->streams[i]);
  });
}

int fa_mgpu_ring_forward(void *group, const void *const *Q, const void *const *K, const void *const *V, void *const *O,
                         float *const *L, int n_local, int D, int H, float scale, int is_causal, int dtype) {
  DeviceGuard guard;
  MgpuGroup *g = reinterpret_cast<MgpuGroup *>(group);
  FA_REQUIRE(g && Q && K && V && O, "null argument");
  const size_t wsb = fa_ring_workspace_bytes_ex(g->n, FA_RING_TRANSPORT_PEER, n_local, D, H, dtype);
  FA_REQUIRE(wsb > 0, "bad shape");
  const size_t tile_bytes = (size_t)H * n_local * D * 2;
  int rc = group_reserve(g, wsb, g->n > 1 ? 2 * tile_bytes : 0);  // window: [K | V]
  if (rc != FA_OK) return rc;
  return run_on_all(g, [&](int i) -> int {
    return fa_ring_attention_forward(g->rings[i], Q[i], K[i], V[i], O[i], L ? L[i] : nullptr, n_local, D, H, scale, is_causal,
                                     dtype, g->ws[i], wsb, g->streams[i]);
  });
}

int fa_mgpu_ring_backward(void *group, const void *const *Q, const void *const *K, const void *const *V,
                          const void *const *O, const void *const *dO, const float *const *L, float *const *dQ,
                          float *const *dK, float *const *dV, int n_local, int D, int H, float scale, int is_causal,
                          int dtype) {
  DeviceGuard guard;
  MgpuGroup *g = reinterpret_cast<MgpuGroup *>(group);
  FA_REQUIRE(g && Q && K && V && O && dO && L && dQ && dK && dV, "null argument");
  const size_t wsb = fa_ring_workspace_bytes_backward(n_local, D, H, dtype);
  FA_REQUIRE(wsb > 0, "bad shape");
  const size_t tile_elems = (size_t)H * n_local * D;
  int rc = group_reserve(g, wsb, g->n > 1 ? 2 * tile_elems * 2 + 4 * tile_elems * 4 : 0);  // window: [K | V][2 x (dK | dV)]
  if (rc != FA_OK) return rc;
  return run_on_all(g, [&](int i) -> int {
    return fa_ring_attention_backward(g->rings[i], Q[i], K[i], V[i], O[i], dO[i], L[i], dQ[i], dK[i], dV[i], n_local, D, H, scale,
                                      is_causal, dtype, g->ws[i], wsb, g->streams[i]);
  });
}

}  // extern "C"
