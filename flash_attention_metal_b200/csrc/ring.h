// State of one rank of a ring-attention group (ring.cu) and of a single-process multi-GPU group
// (mgpu.cu).  Internal to libflash_attn_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace fa {

constexpr int kRingMaxWorld = 16;  // one box: at most 8 GPUs today

struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
struct MgpuGroup;

struct Ring {
  NcclComm comm = nullptr;   // null inside an fa_mgpu group (one process: no NCCL needed)
  MgpuGroup *group = nullptr;
  int rank = 0, world = 1, device = 0;
  int transport = 0;         // FA_RING_TRANSPORT_* in effect (never AUTO)
  unsigned scratch_next = 0;  // next local staging word for flag values (ring.cu flag_write)
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t inputs_ready = nullptr, comm_idle = nullptr;
  cudaEvent_t recv_done[2] = {nullptr, nullptr}, compute_done[2] = {nullptr, nullptr};
  cudaEvent_t add_done[2] = {nullptr, nullptr}, acc_recv_done[2] = {nullptr, nullptr};  // backward dK/dV ring
  // ---- PEER transport ----
  uint32_t *flags = nullptr;                 // this rank's flag window (device memory, written by peers)
  uint32_t *peer_flags[kRingMaxWorld] = {};  // every rank's flag window as mapped here
  char *data = nullptr;                      // this rank's data window: [K | V] [dK|dV accumulator x 2]
  size_t data_cap = 0;
  char *peer_data[kRingMaxWorld] = {};
  uint32_t epoch = 0;                        // K/V publications so far (one per forward / backward call)
  uint32_t acc_pub = 0, acc_pull = 0;        // dK/dV accumulators published / pulled so far
  char *xchg = nullptr;                      // device staging of the handle exchange (multi-process only)
};

int ring_init_streams(Ring *r);
int ring_alloc_flags(Ring *r);
int ring_ensure_window(Ring *r, size_t bytes);
void ring_free(Ring *r);
int mgpu_grow_windows(MgpuGroup *g, size_t bytes);

}  // namespace fa
