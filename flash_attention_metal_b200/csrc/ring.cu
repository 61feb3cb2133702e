// Ring / context-parallel attention across the GPUs of one box (not in the reference:
// BASELINE.json config 5).  Each rank owns n_local query rows of every head and the matching
// K/V rows; at step s it attends its queries to the K/V chunk of rank (rank - s) mod P and folds
// the partial result into a running fp32 (O, L) pair with the online-softmax rule the reference
// uses inside its kernel (kernels.metal:784-791):  L' = log(e^L1 + e^L2),
// O' = O1 e^(L1-L') + O2 e^(L2-L').  The fold happens in the EPILOGUE of the forward kernel
// (FwdMerge, fwd_tc.cu): one launch per ring step, no separate merge kernel, no 16-bit round trip.
//
// Causal balance: zig-zag.  The sequence is cut into 2P chunks of c = n_local/2 rows and
// rank r holds chunks r and 2P-1-r (local rows [0,c) and [c,2c)).  Then at every ring step
// each rank has exactly two c x c blocks of unmasked work (fa_ring_plan):
//   step 0 (own K/V)      : causal attention over the local 2c rows in local order
//   K/V from a lower rank : all 2c local queries x the first c received keys
//   K/V from a higher rank: the last c local queries x all 2c received keys
// and fully masked chunk pairs are never computed, waited for, or (peer transport) transferred.
//
// Transports (chosen once per ring, agreed by all ranks at creation -- never per call, never by
// environment variable):
//   PEER        every rank exposes a window of device memory to its peers (CUDA IPC between
//               processes, plain peer access inside one process).  K/V chunks are PULLED straight
//               from their owner by the copy engines (cudaMemcpyAsync over NVLink; through
//               NVSwitch every owner is one hop away, so nothing is forwarded hop by hop), and
//               ranks synchronise with 32-bit flags in each other's windows driven by stream
//               memory operations (cuStreamWriteValue32 / cuStreamWaitValue32).  No SM is used for
//               communication: the 1-CTA-per-SM compute grids keep the whole GPU.
//   NCCL        ncclSend/ncclRecv ring on a side stream (NCCL's copy kernels compete with the
//               compute grid for SMs: measured 67 GB/s effective at medium N in round 1).
//   NCCL_GATHER one ncclAllGather of all K/V under the local block, remote blocks back to back.
// NCCL is dlopen()ed so that the library has no link-time dependency on it; a single-process
// group (fa_mgpu_*, mgpu.cu) uses the PEER transport without NCCL at all.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "fa_internal.h"
#include "ring.h"

namespace fa {
namespace {

// ---- the handful of NCCL entry points used, resolved at run time ------------------
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId *);
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int);
  int (*CommDestroy)(NcclComm);
  int (*Send)(const void *, size_t, int /*dtype*/, int, NcclComm, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t);
  int (*AllGather)(const void *, void *, size_t, int /*dtype*/, NcclComm, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char *(*GetErrorString)(int);
  void *handle = nullptr;
};
constexpr int kNcclUint8 = 1;

NcclApi *nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define FA_SYM(field, sym) *(void **)(&api.field) = dlsym(api.handle, sym)
    FA_SYM(GetUniqueId, "ncclGetUniqueId");
    FA_SYM(CommInitRank, "ncclCommInitRank");
    FA_SYM(CommDestroy, "ncclCommDestroy");
    FA_SYM(Send, "ncclSend");
    FA_SYM(Recv, "ncclRecv");
    FA_SYM(AllGather, "ncclAllGather");
    FA_SYM(GroupStart, "ncclGroupStart");
    FA_SYM(GroupEnd, "ncclGroupEnd");
    FA_SYM(GetErrorString, "ncclGetErrorString");
#undef FA_SYM
    if (!api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.AllGather || !api.GroupStart ||
        !api.GroupEnd) {
      dlclose(api.handle);
      api.handle = nullptr;
    }
  });
  return api.handle ? &api : nullptr;
}

int nccl_error(const char *what, int code) {
  NcclApi *api = nccl();
  return set_error(FA_ERR_NCCL, "%s failed: %s", what,
                   api && api->GetErrorString ? api->GetErrorString(code) : "nccl error");
}
#define FA_NCCL_CHECK(expr)                      \
  do {                                           \
    int _r = (expr);                             \
    if (_r != 0) return nccl_error(#expr, _r);   \
  } while (0)

// A send/recv group that is always closed, also on the error path (an open group would swallow the
// next NCCL call of this thread).
template <typename F>
int nccl_group(NcclApi *api, F &&body) {
  int rc = api->GroupStart();
  if (rc != 0) return nccl_error("ncclGroupStart", rc);
  const int body_rc = body();
  rc = api->GroupEnd();
  if (body_rc != FA_OK) return body_rc;
  if (rc != 0) return nccl_error("ncclGroupEnd", rc);
  return FA_OK;
}

// flag words of a rank's flag window
constexpr int kFlagKvReady = 0;                 // [src]   src's K/V window holds the data of epoch e
constexpr int kFlagKvDone = kRingMaxWorld;      // [peer]  peer has finished reading my K/V window of epoch e
constexpr int kFlagAccReady = 2 * kRingMaxWorld;      // prev has published its k-th dK/dV accumulator
constexpr int kFlagAccDone = 2 * kRingMaxWorld + 1;   // next has pulled my k-th accumulator
constexpr int kFlagScratch = 256, kFlagScratchWords = 512;  // local staging of flag values (third write mechanism)
constexpr size_t kFlagBytes = 4096;

// ---- stream memory operations (driver API, resolved through the runtime) ----------
typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*MemsetD32AsyncFn)(CUdeviceptr, unsigned int, size_t, CUstream);
struct MemOps {
  StreamValue32Fn wait = nullptr, write = nullptr;
  MemsetD32AsyncFn memset32 = nullptr;
};
const MemOps *memops() {
  static MemOps ops;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char *name) -> void * {
      void *p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        return nullptr;
      }
      return p;
    };
    ops.wait = reinterpret_cast<StreamValue32Fn>(get("cuStreamWaitValue32"));
    ops.write = reinterpret_cast<StreamValue32Fn>(get("cuStreamWriteValue32"));
    ops.memset32 = reinterpret_cast<MemsetD32AsyncFn>(get("cuMemsetD32Async"));
  });
  return &ops;
}

// stream `st` proceeds once *flag >= value (flag is in this device's memory; peers write it)
int flag_wait(cudaStream_t st, const uint32_t *flag, uint32_t value) {
  const MemOps *m = memops();
  if (!m->wait) return set_error(FA_ERR_UNSUPPORTED, "cuStreamWaitValue32 is not available from the driver");
  const CUresult r = m->wait((CUstream)st, (CUdeviceptr)(uintptr_t)flag, value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) return set_error(FA_ERR_CUDA, "cuStreamWaitValue32 failed (CUresult %d)", (int)r);
  return FA_OK;
}
// *flag = value, in stream order after everything enqueued on `st` so far.  The flag usually lives in
// a peer's window.  Three mechanisms, tried in this order and remembered per process: a stream write
// straight to the peer address; a 32-bit memset of the peer address; a stream write into local
// scratch followed by a 4-byte copy to the peer (always possible: it is an ordinary peer copy).
std::atomic<int> g_write_mech{0};
int flag_write(cudaStream_t st, uint32_t *flag, uint32_t value, Ring *r) {
  const MemOps *m = memops();
  int mech = g_write_mech.load(std::memory_order_relaxed);
  if (mech == 0) {
    if (m->write && m->write((CUstream)st, (CUdeviceptr)(uintptr_t)flag, value, 0) == CUDA_SUCCESS) return FA_OK;
    g_write_mech.store(mech = 1, std::memory_order_relaxed);
  }
  if (mech == 1) {
    if (m->memset32 && m->memset32((CUdeviceptr)(uintptr_t)flag, value, 1, (CUstream)st) == CUDA_SUCCESS) return FA_OK;
    g_write_mech.store(mech = 2, std::memory_order_relaxed);
  }
  (void)cudaGetLastError();
  if (!m->write || !r || !r->flags) return set_error(FA_ERR_UNSUPPORTED, "no way to write a peer flag from a stream");
  uint32_t *scratch = r->flags + kFlagScratch + (r->scratch_next++ % kFlagScratchWords);
  const CUresult e = m->write((CUstream)st, (CUdeviceptr)(uintptr_t)scratch, value, 0);
  if (e != CUDA_SUCCESS) return set_error(FA_ERR_CUDA, "cuStreamWriteValue32 failed (CUresult %d)", (int)e);
  FA_CUDA_CHECK(cudaMemcpyAsync(flag, scratch, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  return FA_OK;
}

// Backward dK/dV ring step: out = (has_in ? in : 0) (+ tmp on the key rows this step touched).
// One launch covers dK and dV ([2][H, n_local, D] fp32 each, stacked).  Single owner per element.
__global__ void __launch_bounds__(256) ring_dkv_add_kernel(float *__restrict__ out, const float *__restrict__ in,
                                                           const float *__restrict__ tmp, int64_t per_tensor,
                                                           int n_local, int D, int k_off, int k_rows, int has_in) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index within one tensor
  if (v * 4 >= per_tensor) return;
  const int64_t e = v * 4 + (int64_t)blockIdx.y * per_tensor;        // blockIdx.y: 0 = dK, 1 = dV
  const int row = (int)(((v * 4) / D) % n_local);
  float4 a = has_in ? *reinterpret_cast<const float4 *>(in + e) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= k_off && row < k_off + k_rows) {
    const float4 t = *reinterpret_cast<const float4 *>(tmp + e);
    a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
  }
  *reinterpret_cast<float4 *>(out + e) = a;
}

struct Block { int src, q_off, q_rows, k_off, k_rows, causal; };

// The schedule (pure host logic, also exported as fa_ring_plan and tested on CPU).
int ring_plan(int rank, int world, int step, int n_local, int is_causal, Block *b) {
  if (world < 1 || rank < 0 || rank >= world || step < 0 || step >= world || n_local < 1) return -1;
  b->src = ((rank - step) % world + world) % world;
  if (!is_causal) {
    *b = Block{b->src, 0, n_local, 0, n_local, 0};
    return 0;
  }
  if (n_local % 2) return -1;
  const int c = n_local / 2;
  if (step == 0) *b = Block{b->src, 0, n_local, 0, n_local, 1};        // own chunks: local causal
  else if (b->src < rank) *b = Block{b->src, 0, n_local, 0, c, 0};      // every local query sees chunk src only
  else *b = Block{b->src, c, c, 0, n_local, 0};                         // only chunk 2P-1-rank sees both received chunks
  return 0;
}

int merge_mode(bool first, bool last) {
  return first ? (last ? 0 : 1) : (last ? 3 : 2);  // kMergeNone / First / Last / Middle (fwd_tc.cu)
}

// How the forward launch of ring step `step` folds its partial into the running (O, L) pair (pure host logic,
// exported as fa_ring_merge_plan and tested on CPU): merge modes for the rows of the launch below / from
// `half_rows`.  Not causal: every local row takes part in every step.  Causal (zig-zag): the first chunk (local
// rows < c) takes part in steps 0 .. rank, the second chunk in every step; a launch that covers only the second
// chunk (K/V from a higher rank) has one row range.
void merge_plan(int rank, int world, int step, int n_local, int is_causal, int q_off, int *lo, int *hi, int *half_rows) {
  if (!is_causal) {
    *lo = *hi = merge_mode(step == 0, step == world - 1);
    *half_rows = 0;
    return;
  }
  const int first_chunk = merge_mode(step == 0, step == rank), second_chunk = merge_mode(step == 0, step == world - 1);
  if (q_off == 0) { *lo = first_chunk; *hi = second_chunk; *half_rows = n_local / 2; }
  else { *lo = *hi = second_chunk; *half_rows = 0; }
}


struct IpcBlob {  // what ranks exchange about a window (padded to 128 bytes)
  cudaIpcMemHandle_t handle;
  int ok;
  char pad[128 - sizeof(cudaIpcMemHandle_t) - sizeof(int)];
};
static_assert(sizeof(IpcBlob) == 128, "IpcBlob must be 128 bytes");

// all-gather of one 128-byte blob per rank through NCCL (host in, host out; synchronous)
int exchange_blobs(Ring *r, const IpcBlob &mine, IpcBlob *all) {
  NcclApi *api = nccl();
  if (!api || !r->comm) return set_error(FA_ERR_NCCL, "no NCCL communicator for the handle exchange");
  FA_CUDA_CHECK(cudaMemcpy(r->xchg + (size_t)r->rank * sizeof(IpcBlob), &mine, sizeof(IpcBlob), cudaMemcpyHostToDevice));
  FA_NCCL_CHECK(api->AllGather(r->xchg + (size_t)r->rank * sizeof(IpcBlob), r->xchg, sizeof(IpcBlob), kNcclUint8, r->comm,
                               r->comm_stream));
  FA_CUDA_CHECK(cudaStreamSynchronize(r->comm_stream));
  FA_CUDA_CHECK(cudaMemcpy(all, r->xchg, (size_t)r->world * sizeof(IpcBlob), cudaMemcpyDeviceToHost));
  return FA_OK;
}

// Map every peer's window `mine` belongs to.  Collective.  *all_ok tells whether every rank succeeded
// (the same answer on every rank); on failure nothing stays mapped.
int map_peer_windows(Ring *r, void *mine, void **peer_out, bool *all_ok) {
  IpcBlob blob = {}, all[kRingMaxWorld];
  blob.ok = mine != nullptr && cudaIpcGetMemHandle(&blob.handle, mine) == cudaSuccess;
  (void)cudaGetLastError();
  int rc = exchange_blobs(r, blob, all);
  if (rc != FA_OK) return rc;
  bool ok = true;
  for (int p = 0; p < r->world; ++p) ok = ok && all[p].ok;
  for (int p = 0; p < r->world; ++p) peer_out[p] = nullptr;
  if (ok) {
    for (int p = 0; p < r->world && ok; ++p) {
      if (p == r->rank) { peer_out[p] = mine; continue; }
      if (cudaIpcOpenMemHandle(&peer_out[p], all[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        (void)cudaGetLastError();
        peer_out[p] = nullptr;
        ok = false;
      }
    }
  }
  // second round: did everybody manage to open everything?
  blob.ok = ok;
  rc = exchange_blobs(r, blob, all);
  if (rc != FA_OK) return rc;
  bool everyone = true;
  for (int p = 0; p < r->world; ++p) everyone = everyone && all[p].ok;
  if (!everyone)
    for (int p = 0; p < r->world; ++p) {
      if (p != r->rank && peer_out[p]) cudaIpcCloseMemHandle(peer_out[p]);
      peer_out[p] = nullptr;
    }
  *all_ok = everyone;
  return FA_OK;
}

void unmap_peer_windows(Ring *r, void **peers) {
  for (int p = 0; p < r->world; ++p) {
    if (p != r->rank && peers[p] && !r->group) cudaIpcCloseMemHandle(peers[p]);
    peers[p] = nullptr;
  }
}

}  // namespace

int preload_ring() {
  cudaFuncAttributes attr;
  FA_CUDA_CHECK(cudaFuncGetAttributes(&attr, ring_dkv_add_kernel));
  return FA_OK;
}

// ---- ring life cycle (also used by mgpu.cu) ----------------------------------------
int ring_init_streams(Ring *r) {
  FA_CUDA_CHECK(cudaStreamCreateWithFlags(&r->comm_stream, cudaStreamNonBlocking));
  FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->inputs_ready, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->recv_done[i], cudaEventDisableTiming));
    FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->compute_done[i], cudaEventDisableTiming));
    FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->add_done[i], cudaEventDisableTiming));
    FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->acc_recv_done[i], cudaEventDisableTiming));
  }
  FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->comm_idle, cudaEventDisableTiming));
  return FA_OK;
}

int ring_alloc_flags(Ring *r) {
  FA_CUDA_CHECK(cudaMalloc(&r->flags, kFlagBytes));
  FA_CUDA_CHECK(cudaMemset(r->flags, 0, kFlagBytes));
  return FA_OK;
}

// Make the peer-visible data window at least `bytes` large.  Collective (every rank calls it with the
// same size, so every rank takes the same branch); synchronises the device when it has to grow.
int ring_ensure_window(Ring *r, size_t bytes) {
  if (r->data_cap >= bytes) return FA_OK;
  if (r->group) return mgpu_grow_windows(r->group, bytes);  // one process: the group reallocates every rank's window
  FA_CUDA_CHECK(cudaDeviceSynchronize());
  unmap_peer_windows(r, reinterpret_cast<void **>(r->peer_data));
  // nobody may free a window a peer still has mapped: a round of the exchange is the barrier
  IpcBlob blob = {}, all[kRingMaxWorld];
  int rc = exchange_blobs(r, blob, all);
  if (rc != FA_OK) return rc;
  if (r->data) cudaFree(r->data);
  r->data = nullptr;
  r->data_cap = 0;
  const size_t cap = (bytes + (size_t(1) << 21) - 1) & ~((size_t(1) << 21) - 1);
  void *p = nullptr;
  const bool alloc_ok = cudaMalloc(&p, cap) == cudaSuccess;
  (void)cudaGetLastError();
  bool ok = false;
  rc = map_peer_windows(r, alloc_ok ? p : nullptr, reinterpret_cast<void **>(r->peer_data), &ok);
  if (rc != FA_OK || !ok) {
    if (p) cudaFree(p);
    return rc != FA_OK ? rc : set_error(FA_ERR_CUDA, "could not allocate and share a %zu-byte peer window on every rank", cap);
  }
  r->data = reinterpret_cast<char *>(p);
  r->data_cap = cap;
  return FA_OK;
}

void ring_free(Ring *r) {
  if (!r) return;
  DeviceGuard guard;
  cudaSetDevice(r->device);
  cudaDeviceSynchronize();
  if (!r->group) {
    unmap_peer_windows(r, reinterpret_cast<void **>(r->peer_data));
    unmap_peer_windows(r, reinterpret_cast<void **>(r->peer_flags));
  }
  if (r->comm && nccl() && nccl()->CommDestroy) nccl()->CommDestroy(r->comm);
  if (r->data) cudaFree(r->data);
  if (r->flags) cudaFree(r->flags);
  if (r->xchg) cudaFree(r->xchg);
  if (r->comm_stream) cudaStreamDestroy(r->comm_stream);
  if (r->inputs_ready) cudaEventDestroy(r->inputs_ready);
  if (r->comm_idle) cudaEventDestroy(r->comm_idle);
  for (int i = 0; i < 2; ++i) {
    if (r->recv_done[i]) cudaEventDestroy(r->recv_done[i]);
    if (r->compute_done[i]) cudaEventDestroy(r->compute_done[i]);
    if (r->add_done[i]) cudaEventDestroy(r->add_done[i]);
    if (r->acc_recv_done[i]) cudaEventDestroy(r->acc_recv_done[i]);
  }
  (void)cudaGetLastError();
  delete r;
}

namespace {

// K/V chunk rows [k_off, k_off + k_rows) of every head: contiguous when the whole chunk is wanted,
// otherwise a strided 2-D copy (both run on the copy engines)
int copy_rows(void *dst, const void *src, int n_local, int D, int H, int k_off, int k_rows, cudaStream_t st) {
  const size_t row = (size_t)D * 2;
  if (k_off == 0 && k_rows == n_local) {
    FA_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)H * n_local * row, cudaMemcpyDeviceToDevice, st));
  } else {
    FA_CUDA_CHECK(cudaMemcpy2DAsync(reinterpret_cast<char *>(dst) + k_off * row, (size_t)n_local * row,
                                    reinterpret_cast<const char *>(src) + k_off * row, (size_t)n_local * row,
                                    (size_t)k_rows * row, H, cudaMemcpyDeviceToDevice, st));
  }
  return FA_OK;
}

// PEER transport, start of a call: publish this rank's K/V in its window (epoch e) and tell every peer.
int peer_publish_kv(Ring *r, const void *K, const void *V, size_t tile_bytes, uint32_t e) {
  cudaStream_t cs = r->comm_stream;
  for (int p = 0; p < r->world; ++p)  // every peer has finished reading the previous epoch
    if (p != r->rank) {
      int rc = flag_wait(cs, r->flags + kFlagKvDone + p, e - 1);
      if (rc != FA_OK) return rc;
    }
  FA_CUDA_CHECK(cudaMemcpyAsync(r->data, K, tile_bytes, cudaMemcpyDeviceToDevice, cs));
  FA_CUDA_CHECK(cudaMemcpyAsync(r->data + tile_bytes, V, tile_bytes, cudaMemcpyDeviceToDevice, cs));
  for (int p = 0; p < r->world; ++p)
    if (p != r->rank) {
      int rc = flag_write(cs, r->peer_flags[p] + kFlagKvReady + r->rank, e, r);
      if (rc != FA_OK) return rc;
    }
  return FA_OK;
}

// PEER transport: pull the part of rank src's K/V that block b needs into `slot` ([K | V], chunk layout)
int peer_pull_kv(Ring *r, const Block &b, char *slot, size_t tile_bytes, int n_local, int D, int H, uint32_t e) {
  cudaStream_t cs = r->comm_stream;
  int rc = flag_wait(cs, r->flags + kFlagKvReady + b.src, e);
  if (rc != FA_OK) return rc;
  const char *src = r->peer_data[b.src];
  if ((rc = copy_rows(slot, src, n_local, D, H, b.k_off, b.k_rows, cs)) != FA_OK) return rc;
  if ((rc = copy_rows(slot + tile_bytes, src + tile_bytes, n_local, D, H, b.k_off, b.k_rows, cs)) != FA_OK) return rc;
  return flag_write(cs, r->peer_flags[b.src] + kFlagKvDone + r->rank, e, r);
}

}  // namespace
}  // namespace fa

using namespace fa;

extern "C" {

// Development aids (not in the public header): which of the three peer-flag write mechanisms this process
// settled on (0 stream write to the peer address, 1 32-bit memset of it, 2 local stream write + 4-byte
// peer copy), and a way to force one.
int fa_debug_ring_write_mechanism(void) { return g_write_mech.load(std::memory_order_relaxed); }
void fa_debug_set_ring_write_mechanism(int m) { g_write_mech.store(m < 0 ? 0 : (m > 2 ? 2 : m), std::memory_order_relaxed); }
// copy of the first n flag words of a ring (synchronous)
int fa_debug_ring_flags(void *ring, uint32_t *out, int n) {
  Ring *r = reinterpret_cast<Ring *>(ring);
  FA_REQUIRE(r && r->flags && out && n > 0 && n <= 1024, "bad arguments");
  DeviceGuard guard;
  FA_CUDA_CHECK(cudaSetDevice(r->device));
  FA_CUDA_CHECK(cudaMemcpy(out, r->flags, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return FA_OK;
}

int fa_ring_unique_id_bytes(void) { return (int)sizeof(NcclUniqueId); }

int fa_ring_get_unique_id(void *out, int bytes) {
  FA_REQUIRE(out && bytes >= (int)sizeof(NcclUniqueId), "unique-id buffer must hold %d bytes", (int)sizeof(NcclUniqueId));
  NcclApi *api = nccl();
  if (!api) return set_error(FA_ERR_NCCL, "libnccl.so.2 could not be loaded");
  NcclUniqueId id;
  FA_NCCL_CHECK(api->GetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return FA_OK;
}

int fa_ring_create_ex(void **ring_out, const void *unique_id, int rank, int world, int device, int transport) {
  FA_REQUIRE(ring_out && unique_id && world >= 1 && world <= kRingMaxWorld && rank >= 0 && rank < world,
             "bad ring arguments (world must be 1..%d)", kRingMaxWorld);
  FA_REQUIRE(transport >= FA_RING_TRANSPORT_AUTO && transport <= FA_RING_TRANSPORT_PEER, "unknown ring transport %d", transport);
  NcclApi *api = nccl();
  if (!api) return set_error(FA_ERR_NCCL, "libnccl.so.2 could not be loaded");
  DeviceGuard guard;
  FA_CUDA_CHECK(cudaSetDevice(device));
  Ring *r = new Ring();
  r->rank = rank; r->world = world; r->device = device;
  auto fail = [&](int rc) { ring_free(r); return rc; };
  NcclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  int nrc = api->CommInitRank(&r->comm, world, id, rank);
  if (nrc != 0) {
    r->comm = nullptr;
    return fail(nccl_error("ncclCommInitRank", nrc));
  }
  int rc = ring_init_streams(r);
  if (rc != FA_OK) return fail(rc);
  if (cudaMalloc(&r->xchg, sizeof(IpcBlob) * kRingMaxWorld) != cudaSuccess)
    return fail(set_error(FA_ERR_CUDA, "cudaMalloc of the exchange buffer failed"));
  r->transport = transport == FA_RING_TRANSPORT_AUTO ? FA_RING_TRANSPORT_NCCL : transport;
  if (world > 1 && (transport == FA_RING_TRANSPORT_AUTO || transport == FA_RING_TRANSPORT_PEER)) {
    // try to set the peer transport up; the outcome is agreed by all ranks inside map_peer_windows
    const MemOps *m = memops();
    const bool have_ops = m->wait && m->write;
    bool ok = false;
    if (have_ops) (void)ring_alloc_flags(r);
    rc = map_peer_windows(r, have_ops ? r->flags : nullptr, reinterpret_cast<void **>(r->peer_flags), &ok);
    if (rc != FA_OK) return fail(rc);
    if (ok) {
      r->transport = FA_RING_TRANSPORT_PEER;
    } else if (transport == FA_RING_TRANSPORT_PEER) {
      return fail(set_error(FA_ERR_UNSUPPORTED, "the peer transport (CUDA IPC + stream memory operations) is not available on every rank"));
    }
  }
  *ring_out = r;
  return FA_OK;
}

int fa_ring_create(void **ring_out, const void *unique_id, int rank, int world, int device) {
  return fa_ring_create_ex(ring_out, unique_id, rank, world, device, FA_RING_TRANSPORT_AUTO);
}

int fa_ring_transport(void *ring) {
  Ring *r = reinterpret_cast<Ring *>(ring);
  FA_REQUIRE(r, "null ring");
  return r->transport;
}

int fa_ring_destroy(void *ring) {
  Ring *r = reinterpret_cast<Ring *>(ring);
  if (!r) return FA_OK;
  if (r->group) return set_error(FA_ERR_INVALID, "this ring belongs to an fa_mgpu group: destroy the group");
  ring_free(r);
  return FA_OK;
}

int fa_ring_plan(int rank, int world, int step, int n_local, int is_causal, int *src_rank, int *q_off,
                 int *q_rows, int *k_off, int *k_rows, int *block_causal) {
  Block b;
  if (ring_plan(rank, world, step, n_local, is_causal, &b) != 0)
    return set_error(FA_ERR_INVALID, "bad ring plan arguments (rank %d world %d step %d n_local %d causal %d)", rank,
                     world, step, n_local, is_causal);
  *src_rank = b.src; *q_off = b.q_off; *q_rows = b.q_rows; *k_off = b.k_off; *k_rows = b.k_rows; *block_causal = b.causal;
  return FA_OK;
}

// Merge modes of the forward launch of a ring step (0 none, 1 first, 2 middle, 3 last) for its rows below /
// from half_rows (row indices relative to the launch's first query row).  Host-only.
int fa_ring_merge_plan(int rank, int world, int step, int n_local, int is_causal, int *lo, int *hi, int *half_rows) {
  Block b;
  if (!lo || !hi || !half_rows || ring_plan(rank, world, step, n_local, is_causal, &b) != 0)
    return set_error(FA_ERR_INVALID, "bad ring plan arguments (rank %d world %d step %d n_local %d causal %d)", rank,
                     world, step, n_local, is_causal);
  merge_plan(rank, world, step, n_local, is_causal, b.q_off, lo, hi, half_rows);
  return FA_OK;
}

// Global row index of the first row of each local chunk (chunk 1 has 0 rows when not causal).
int fa_ring_local_rows(int rank, int world, int n_local, int is_causal, int64_t first_row[2], int rows[2]) {
  FA_REQUIRE(world >= 1 && rank >= 0 && rank < world && n_local >= 1, "bad arguments");
  if (!is_causal) {
    first_row[0] = (int64_t)rank * n_local; rows[0] = n_local;
    first_row[1] = 0; rows[1] = 0;
    return FA_OK;
  }
  FA_REQUIRE(n_local % 2 == 0, "causal ring attention needs an even n_local (two zig-zag chunks)");
  const int c = n_local / 2;
  first_row[0] = (int64_t)rank * c; rows[0] = c;
  first_row[1] = (int64_t)(2 * world - 1 - rank) * c; rows[1] = c;
  return FA_OK;
}

// Forward workspace: [receive slots: 2 x (K | V), or world x (K | V) for NCCL_GATHER] [O_acc fp32] [L_acc fp32]
size_t fa_ring_workspace_bytes_ex(int world, int transport, int n_local, int D, int H, int dtype) {
  (void)dtype;
  if (n_local < 1 || H < 1 || D < 1 || world < 1) return 0;
  const size_t tile = (size_t)H * n_local * D;
  const size_t slots = transport == FA_RING_TRANSPORT_NCCL_GATHER ? (size_t)world : 2;
  return slots * (2 * tile * 2) + tile * 4 + (size_t)H * n_local * 4 + 1024;
}
size_t fa_ring_workspace_bytes(int n_local, int D, int H, int dtype) {
  return fa_ring_workspace_bytes_ex(2, FA_RING_TRANSPORT_AUTO, n_local, D, H, dtype);
}
size_t fa_ring_workspace_bytes_gather(int world, int n_local, int D, int H, int dtype) {
  return fa_ring_workspace_bytes_ex(world, FA_RING_TRANSPORT_NCCL_GATHER, n_local, D, H, dtype);
}

int fa_ring_attention_forward(void *ring, const void *Q, const void *K, const void *V, void *O, float *L_out,
                              int n_local, int D, int H, float scale, int is_causal, int dtype, void *workspace,
                              size_t workspace_bytes, fa_stream_t stream_) {
  Ring *r = reinterpret_cast<Ring *>(ring);
  FA_REQUIRE(r && Q && K && V && O, "null argument");
  FA_REQUIRE(D == 64 || D == 128, "D must be 64 or 128 (got %d)", D);
  FA_REQUIRE(!is_causal || n_local % 2 == 0, "causal ring attention needs an even n_local");
  FA_REQUIRE(n_local % 8 == 0, "n_local must be a multiple of 8");
  const int P = r->world;
  const size_t need = fa_ring_workspace_bytes_ex(P, r->transport, n_local, D, H, dtype);
  if (!workspace || workspace_bytes < need)
    return set_error(FA_ERR_WORKSPACE, "ring workspace too small for transport %d: need %zu bytes, got %zu",
                     r->transport, need, workspace_bytes);
  NcclApi *api = nccl();
  if (P > 1 && r->transport != FA_RING_TRANSPORT_PEER && !api) return set_error(FA_ERR_NCCL, "libnccl.so.2 could not be loaded");
  DeviceGuard guard;
  FA_CUDA_CHECK(cudaSetDevice(r->device));
  cudaStream_t st = (cudaStream_t)stream_;
  const size_t tile_elems = (size_t)H * n_local * D;
  const size_t tile_bytes = tile_elems * 2;
  char *ws = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  const bool gather = r->transport == FA_RING_TRANSPORT_NCCL_GATHER && P > 1;
  const size_t kv_area = (gather ? (size_t)P : 2) * 2 * tile_bytes;
  char *slot[2] = {ws, ws + 2 * tile_bytes};
  char *k_all = ws, *v_all = ws + (size_t)P * tile_bytes;
  float *o_acc = reinterpret_cast<float *>(ws + kv_area);
  float *l_acc = reinterpret_cast<float *>(ws + kv_area + tile_elems * 4);
  const int next = (r->rank + 1) % P, prev = (r->rank - 1 + P) % P;
  const int64_t hs = (int64_t)n_local * D;

  // ---- local work of ring step s on the chunk (curK, curV) of rank (rank - s) mod P: ONE launch,
  //      the running (O, L) is folded in the kernel epilogue ----
  auto do_step = [&](int s, const void *curK, const void *curV) -> int {
    Block b;
    ring_plan(r->rank, P, s, n_local, is_causal, &b);
    const uint16_t *q = reinterpret_cast<const uint16_t *>(Q) + (int64_t)b.q_off * D;
    const uint16_t *k = reinterpret_cast<const uint16_t *>(curK) + (int64_t)b.k_off * D;
    const uint16_t *v = reinterpret_cast<const uint16_t *>(curV) + (int64_t)b.k_off * D;
    FwdMerge m;
    m.O_acc = o_acc + (int64_t)b.q_off * D;
    m.L_acc = l_acc + b.q_off;
    merge_plan(r->rank, P, s, n_local, is_causal, b.q_off, &m.lo, &m.hi, &m.half_rows);
    return launch_fwd_tc_rect(q, k, v, reinterpret_cast<uint16_t *>(O) + (int64_t)b.q_off * D,
                              L_out ? L_out + b.q_off : nullptr, b.q_rows, b.k_rows, D, scale, (int64_t)H * hs, hs,
                              (int64_t)H * hs, hs, b.causal, 1, H, dtype, st, &m);
  };
  if (P == 1) return do_step(0, K, V);

  FA_CUDA_CHECK(cudaEventRecord(r->inputs_ready, st));
  // the comm stream's work of the previous call on this ring is ordered before this call's
  FA_CUDA_CHECK(cudaStreamWaitEvent(r->comm_stream, r->inputs_ready, 0));

  if (r->transport == FA_RING_TRANSPORT_PEER) {
    int rc = ring_ensure_window(r, 2 * tile_bytes);
    if (rc != FA_OK) return rc;
    const uint32_t e = ++r->epoch;
    if ((rc = peer_publish_kv(r, K, V, tile_bytes, e)) != FA_OK) return rc;
    for (int s = 0; s < P; ++s) {
      if (s + 1 < P) {  // fetch the chunk of step s + 1 while step s computes
        Block nb;
        ring_plan(r->rank, P, s + 1, n_local, is_causal, &nb);
        // slot (s+1)&1 was read by step s-1
        if (s >= 1) FA_CUDA_CHECK(cudaStreamWaitEvent(r->comm_stream, r->compute_done[(s - 1) & 1], 0));
        if ((rc = peer_pull_kv(r, nb, slot[(s + 1) & 1], tile_bytes, n_local, D, H, e)) != FA_OK) return rc;
        FA_CUDA_CHECK(cudaEventRecord(r->recv_done[(s + 1) & 1], r->comm_stream));
      }
      if (s > 0) FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->recv_done[s & 1], 0));
      rc = s == 0 ? do_step(0, K, V) : do_step(s, slot[s & 1], slot[s & 1] + tile_bytes);
      if (rc != FA_OK) return rc;
      FA_CUDA_CHECK(cudaEventRecord(r->compute_done[s & 1], st));
    }
    return FA_OK;
  }

  if (gather) {
    int rc = nccl_group(api, [&]() -> int {
      FA_NCCL_CHECK(api->AllGather(K, k_all, tile_bytes, kNcclUint8, r->comm, r->comm_stream));
      FA_NCCL_CHECK(api->AllGather(V, v_all, tile_bytes, kNcclUint8, r->comm, r->comm_stream));
      return FA_OK;
    });
    if (rc != FA_OK) return rc;
    FA_CUDA_CHECK(cudaEventRecord(r->recv_done[0], r->comm_stream));
    rc = do_step(0, K, V);  // the local block runs under the gather
    if (rc != FA_OK) return rc;
    FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->recv_done[0], 0));
    for (int s = 1; s < P; ++s) {
      const int src = ((r->rank - s) % P + P) % P;
      rc = do_step(s, k_all + (size_t)src * tile_bytes, v_all + (size_t)src * tile_bytes);
      if (rc != FA_OK) return rc;
    }
    return FA_OK;
  }

  // ---- NCCL send/recv ring: the chunk we hold travels on to the next rank while we work on it ----
  const void *curK = K, *curV = V;
  for (int s = 0; s < P; ++s) {
    if (s + 1 < P) {
      // the slot we are about to overwrite was the chunk step s-1 computed on (s >= 2 only)
      if (s >= 2) FA_CUDA_CHECK(cudaStreamWaitEvent(r->comm_stream, r->compute_done[(s - 1) & 1], 0));
      char *dst = slot[s & 1];
      int rc = nccl_group(api, [&]() -> int {
        FA_NCCL_CHECK(api->Send(curK, tile_bytes, kNcclUint8, next, r->comm, r->comm_stream));
        FA_NCCL_CHECK(api->Send(curV, tile_bytes, kNcclUint8, next, r->comm, r->comm_stream));
        FA_NCCL_CHECK(api->Recv(dst, tile_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
        FA_NCCL_CHECK(api->Recv(dst + tile_bytes, tile_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
        return FA_OK;
      });
      if (rc != FA_OK) return rc;
      FA_CUDA_CHECK(cudaEventRecord(r->recv_done[s & 1], r->comm_stream));
    }
    int rc = do_step(s, curK, curV);
    if (rc != FA_OK) return rc;
    FA_CUDA_CHECK(cudaEventRecord(r->compute_done[s & 1], st));
    if (s + 1 < P) {
      FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->recv_done[s & 1], 0));
      curK = slot[s & 1];
      curV = slot[s & 1] + tile_bytes;
    }
  }
  return FA_OK;
}

// Backward workspace.  NCCL: [2 K|V slots][tmp dK|dV][2 outgoing accumulators][incoming accumulator][delta];
// PEER: the outgoing accumulators live in the peer-visible window instead.
size_t fa_ring_workspace_bytes_backward(int n_local, int D, int H, int dtype) {
  (void)dtype;
  if (n_local < 1 || H < 1 || D < 1) return 0;
  const size_t tile = (size_t)H * n_local * D;
  size_t bytes = 2 * (2 * tile * 2)        // two K|V receive slots (16-bit)
                 + 2 * tile * 4            // this step's dK|dV block (fp32)
                 + 2 * (2 * tile * 4)      // two outgoing dK|dV accumulators (NCCL transport)
                 + 2 * tile * 4            // incoming dK|dV accumulator
                 + (size_t)H * n_local * 4   // delta
                 + 256 + bwd_fused_sem_bytes(n_local, 1, H);  // ordering counters of the fused backward kernel
  return bytes + 1024;
}

// Ring backward: K/V chunks travel one step ahead of the tile loop (as in the forward); the dK/dV
// accumulator of a chunk travels one step behind it -- each rank adds the contribution of its own
// queries and passes the sum on, and after the last step one more hop returns the finished dK/dV to
// the owner.  dQ accumulates in place on the owning rank.  Every sum has a fixed order: results are
// run-to-run deterministic.  L must be the final log-sum-exp written by fa_ring_attention_forward.
int fa_ring_attention_backward(void *ring, const void *Q, const void *K, const void *V, const void *O,
                               const void *dO, const float *L, float *dQ, float *dK, float *dV, int n_local, int D,
                               int H, float scale, int is_causal, int dtype, void *workspace, size_t workspace_bytes,
                               fa_stream_t stream_) {
  Ring *r = reinterpret_cast<Ring *>(ring);
  FA_REQUIRE(r && Q && K && V && O && dO && L && dQ && dK && dV, "null argument");
  FA_REQUIRE(D == 64 || D == 128, "D must be 64 or 128 (got %d)", D);
  FA_REQUIRE(!is_causal || n_local % 2 == 0, "causal ring attention needs an even n_local");
  FA_REQUIRE(n_local % 8 == 0, "n_local must be a multiple of 8");
  const size_t need = fa_ring_workspace_bytes_backward(n_local, D, H, dtype);
  if (!workspace || workspace_bytes < need)
    return set_error(FA_ERR_WORKSPACE, "ring backward workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  const int P = r->world;
  const bool peer = r->transport == FA_RING_TRANSPORT_PEER && P > 1;
  NcclApi *api = nccl();
  if (P > 1 && !peer && !api) return set_error(FA_ERR_NCCL, "libnccl.so.2 could not be loaded");
  DeviceGuard guard;
  FA_CUDA_CHECK(cudaSetDevice(r->device));
  cudaStream_t st = (cudaStream_t)stream_, cs = r->comm_stream;
  const size_t tile_elems = (size_t)H * n_local * D;
  const size_t kv_bytes = tile_elems * 2, g_bytes = tile_elems * 4;
  char *ws = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  char *slot[2] = {ws, ws + 2 * kv_bytes};
  float *tmp = reinterpret_cast<float *>(ws + 4 * kv_bytes);                  // [dK | dV] of this step's block
  float *acc_out[2] = {tmp + 2 * tile_elems, tmp + 4 * tile_elems};
  float *acc_in = tmp + 6 * tile_elems;
  float *delta = tmp + 8 * tile_elems;
  void *sems = reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(delta + (size_t)H * n_local) + 255) & ~(uintptr_t)255);
  const int next = (r->rank + 1) % P, prev = (r->rank - 1 + P) % P;
  const int64_t hs = (int64_t)n_local * D;
  int rc = launch_bwd_delta(O, dO, delta, n_local, D, (int64_t)H * hs, hs, 1, H, dtype, st);
  if (rc != FA_OK) return rc;

  auto block_backward = [&](int s, const void *curK, const void *curV, Block *b_out) -> int {
    Block b;
    ring_plan(r->rank, P, s, n_local, is_causal, &b);
    *b_out = b;
    const uint16_t *q = reinterpret_cast<const uint16_t *>(Q) + (int64_t)b.q_off * D;
    const uint16_t *g = reinterpret_cast<const uint16_t *>(dO) + (int64_t)b.q_off * D;
    const uint16_t *k = reinterpret_cast<const uint16_t *>(curK) + (int64_t)b.k_off * D;
    const uint16_t *v = reinterpret_cast<const uint16_t *>(curV) + (int64_t)b.k_off * D;
    float *blk_dk = (P == 1 ? dK : tmp) + (int64_t)b.k_off * D;
    float *blk_dv = (P == 1 ? dV : tmp + tile_elems) + (int64_t)b.k_off * D;
    return launch_bwd_tc_rect(q, k, v, g, L + b.q_off, delta + b.q_off, dQ + (int64_t)b.q_off * D, blk_dk, blk_dv, b.q_rows,
                              b.k_rows, D, scale, (int64_t)H * hs, hs, (int64_t)H * hs, hs, b.causal, s > 0, 1, H, dtype, st, sems);
  };
  auto add_block = [&](float *out, const Block &b, bool has_in) -> int {
    const int64_t vecs = (int64_t)tile_elems / 4;
    dim3 grid((unsigned)((vecs + 255) / 256), 2);
    ring_dkv_add_kernel<<<grid, 256, 0, st>>>(out, acc_in, tmp, (int64_t)tile_elems, n_local, D, b.k_off, b.k_rows, has_in);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return FA_OK;
  };
  if (P == 1) {
    Block b;
    return block_backward(0, K, V, &b);
  }
  FA_CUDA_CHECK(cudaEventRecord(r->inputs_ready, st));
  FA_CUDA_CHECK(cudaStreamWaitEvent(cs, r->inputs_ready, 0));

  if (peer) {
    // window: [K | V] [acc_out 0: dK | dV] [acc_out 1: dK | dV]
    if ((rc = ring_ensure_window(r, 2 * kv_bytes + 4 * g_bytes)) != FA_OK) return rc;
    const uint32_t e = ++r->epoch;
    const size_t acc_off[2] = {2 * kv_bytes, 2 * kv_bytes + 2 * g_bytes};
    float *wacc[2] = {reinterpret_cast<float *>(r->data + acc_off[0]), reinterpret_cast<float *>(r->data + acc_off[1])};
    if ((rc = peer_publish_kv(r, K, V, kv_bytes, e)) != FA_OK) return rc;
    Block nb;
    ring_plan(r->rank, P, 1, n_local, is_causal, &nb);
    if ((rc = peer_pull_kv(r, nb, slot[1], kv_bytes, n_local, D, H, e)) != FA_OK) return rc;
    FA_CUDA_CHECK(cudaEventRecord(r->recv_done[1], cs));
    for (int s = 0; s < P; ++s) {
      Block b;
      if (s > 0) FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->recv_done[s & 1], 0));
      rc = s == 0 ? block_backward(0, K, V, &b) : block_backward(s, slot[s & 1], slot[s & 1] + kv_bytes, &b);
      if (rc != FA_OK) return rc;
      FA_CUDA_CHECK(cudaEventRecord(r->compute_done[s & 1], st));
      // ---- dK/dV accumulator of the chunk we hold: incoming sum (pulled from prev) + our block ----
      if (s > 0) FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->acc_recv_done[s & 1], 0));
      const uint32_t k_pub = r->acc_pub + 1;  // this publication's sequence number
      if (k_pub > 2 && (rc = flag_wait(st, r->flags + kFlagAccDone, k_pub - 2)) != FA_OK) return rc;  // slot reuse
      if ((rc = add_block(wacc[k_pub & 1], b, s > 0)) != FA_OK) return rc;
      FA_CUDA_CHECK(cudaEventRecord(r->add_done[s & 1], st));
      if ((rc = flag_write(st, r->peer_flags[next] + kFlagAccReady, k_pub, r)) != FA_OK) return rc;
      r->acc_pub = k_pub;
      // ---- comm stream: pull what prev published at its step s (its k-th publication, k = ours) ----
      const uint32_t k_pull = ++r->acc_pull;
      if ((rc = flag_wait(cs, r->flags + kFlagAccReady, k_pull)) != FA_OK) return rc;
      const char *src_acc = r->peer_data[prev] + acc_off[k_pull & 1];
      if (s + 1 < P) {
        FA_CUDA_CHECK(cudaStreamWaitEvent(cs, r->add_done[s & 1], 0));  // acc_in was read by this step's add
        FA_CUDA_CHECK(cudaMemcpyAsync(acc_in, src_acc, 2 * g_bytes, cudaMemcpyDeviceToDevice, cs));
      } else {  // last hop: the finished gradients of our own chunk come home
        FA_CUDA_CHECK(cudaMemcpyAsync(dK, src_acc, g_bytes, cudaMemcpyDeviceToDevice, cs));
        FA_CUDA_CHECK(cudaMemcpyAsync(dV, src_acc + g_bytes, g_bytes, cudaMemcpyDeviceToDevice, cs));
      }
      if ((rc = flag_write(cs, r->peer_flags[prev] + kFlagAccDone, k_pull, r)) != FA_OK) return rc;
      FA_CUDA_CHECK(cudaEventRecord(r->acc_recv_done[(s + 1) & 1], cs));
      // ---- K/V of step s + 2 into the slot step s has just finished with ----
      if (s + 2 < P) {
        ring_plan(r->rank, P, s + 2, n_local, is_causal, &nb);
        FA_CUDA_CHECK(cudaStreamWaitEvent(cs, r->compute_done[s & 1], 0));
        if ((rc = peer_pull_kv(r, nb, slot[s & 1], kv_bytes, n_local, D, H, e)) != FA_OK) return rc;
        FA_CUDA_CHECK(cudaEventRecord(r->recv_done[s & 1], cs));
      }
    }
    FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->acc_recv_done[P & 1], 0));
    return FA_OK;
  }

  // ---- NCCL transport ----
  const void *curK = K, *curV = V;
  {  // A(0): our own chunk starts travelling
    rc = nccl_group(api, [&]() -> int {
      FA_NCCL_CHECK(api->Send(curK, kv_bytes, kNcclUint8, next, r->comm, cs));
      FA_NCCL_CHECK(api->Send(curV, kv_bytes, kNcclUint8, next, r->comm, cs));
      FA_NCCL_CHECK(api->Recv(slot[0], kv_bytes, kNcclUint8, prev, r->comm, cs));
      FA_NCCL_CHECK(api->Recv(slot[0] + kv_bytes, kv_bytes, kNcclUint8, prev, r->comm, cs));
      return FA_OK;
    });
    if (rc != FA_OK) return rc;
    FA_CUDA_CHECK(cudaEventRecord(r->recv_done[0], cs));
  }
  for (int s = 0; s < P; ++s) {
    Block b;
    if ((rc = block_backward(s, curK, curV, &b)) != FA_OK) return rc;
    // ---- dK/dV accumulator of the chunk we hold: add our block, pass it on ----
    if (s > 0) FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->acc_recv_done[s & 1], 0));
    if ((rc = add_block(acc_out[s & 1], b, s > 0)) != FA_OK) return rc;
    FA_CUDA_CHECK(cudaEventRecord(r->add_done[s & 1], st));
    FA_CUDA_CHECK(cudaStreamWaitEvent(cs, r->add_done[s & 1], 0));
    rc = nccl_group(api, [&]() -> int {
      FA_NCCL_CHECK(api->Send(acc_out[s & 1], g_bytes, kNcclUint8, next, r->comm, cs));
      FA_NCCL_CHECK(api->Send(acc_out[s & 1] + tile_elems, g_bytes, kNcclUint8, next, r->comm, cs));
      if (s + 1 < P) {
        FA_NCCL_CHECK(api->Recv(acc_in, g_bytes, kNcclUint8, prev, r->comm, cs));
        FA_NCCL_CHECK(api->Recv(acc_in + tile_elems, g_bytes, kNcclUint8, prev, r->comm, cs));
      } else {  // last hop: the finished gradients of our own chunk come home
        FA_NCCL_CHECK(api->Recv(dK, g_bytes, kNcclUint8, prev, r->comm, cs));
        FA_NCCL_CHECK(api->Recv(dV, g_bytes, kNcclUint8, prev, r->comm, cs));
      }
      return FA_OK;
    });
    if (rc != FA_OK) return rc;
    FA_CUDA_CHECK(cudaEventRecord(r->acc_recv_done[(s + 1) & 1], cs));
    // ---- K/V: switch to the chunk that arrived, forward it if somebody still needs it ----
    if (s + 1 < P) {
      FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->recv_done[s & 1], 0));
      curK = slot[s & 1];
      curV = slot[s & 1] + kv_bytes;
      if (s + 2 < P) {  // the other slot was read by step s, which the comm stream has already waited for
        char *dst = slot[(s + 1) & 1];
        rc = nccl_group(api, [&]() -> int {
          FA_NCCL_CHECK(api->Send(curK, kv_bytes, kNcclUint8, next, r->comm, cs));
          FA_NCCL_CHECK(api->Send(curV, kv_bytes, kNcclUint8, next, r->comm, cs));
          FA_NCCL_CHECK(api->Recv(dst, kv_bytes, kNcclUint8, prev, r->comm, cs));
          FA_NCCL_CHECK(api->Recv(dst + kv_bytes, kv_bytes, kNcclUint8, prev, r->comm, cs));
          return FA_OK;
        });
        if (rc != FA_OK) return rc;
        FA_CUDA_CHECK(cudaEventRecord(r->recv_done[(s + 1) & 1], cs));
      }
    }
  }
  FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->acc_recv_done[P & 1], 0));
  return FA_OK;
}

}  // extern "C"
