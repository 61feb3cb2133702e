// Ring / context-parallel attention forward across the GPUs of one box (not in the
// reference: BASELINE.json config 5).  One process per GPU; each rank owns n_local query
// rows of every head and the matching K/V rows.  The K/V chunk travels around the ring with
// NCCL point-to-point calls (ncclSend/ncclRecv over NVLink) on a side stream, double
// buffered, one step ahead of the tile loop that consumes it; partial (O, L) results are
// merged with the online-softmax rule the reference uses inside its kernel
// (kernels.metal:784-791):  L' = log(e^L1 + e^L2),  O' = O1 e^(L1-L') + O2 e^(L2-L').
//
// Causal balance: zig-zag.  The sequence is cut into 2P chunks of c = n_local/2 rows and
// rank r holds chunks r and 2P-1-r (local rows [0,c) and [c,2c)).  Then at every ring step
// each rank has exactly two c x c blocks of unmasked work (fa_ring_plan):
//   step 0 (own K/V)      : causal attention over the local 2c rows in local order
//   K/V from a lower rank : all 2c local queries x the first c received keys
//   K/V from a higher rank: the last c local queries x all 2c received keys
// and fully masked chunk pairs are never computed or waited for.
//
// NCCL is dlopen()ed so that the library has no link-time dependency on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include "fa_internal.h"

namespace fa {
namespace {

// ---- the handful of NCCL entry points used, resolved at run time ------------------
struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId *);
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int);
  int (*CommDestroy)(NcclComm);
  int (*Send)(const void *, size_t, int /*dtype*/, int, NcclComm, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t);
  int (*AllGather)(const void *, void *, size_t, int /*dtype*/, NcclComm, cudaStream_t);  // optional
  int (*GroupStart)();
  int (*GroupEnd)();
  const char *(*GetErrorString)(int);
  void *handle = nullptr;
};
constexpr int kNcclUint8 = 1;

NcclApi *nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
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
      if (!api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd) {
        dlclose(api.handle);
        api.handle = nullptr;
      }
    }
  }
  return api.handle ? &api : nullptr;
}

#define FA_NCCL_CHECK(expr)                                                                      \
  do {                                                                                           \
    int _r = (expr);                                                                             \
    if (_r != 0)                                                                                 \
      return set_error(FA_ERR_NCCL, "%s failed: %s", #expr,                                      \
                       nccl()->GetErrorString ? nccl()->GetErrorString(_r) : "nccl error");      \
  } while (0)

struct Ring {
  NcclComm comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t inputs_ready = nullptr, recv_done[2] = {nullptr, nullptr}, compute_done[2] = {nullptr, nullptr};
  cudaEvent_t add_done[2] = {nullptr, nullptr}, acc_recv_done[2] = {nullptr, nullptr};  // backward dK/dV ring
};

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

// ---- merge kernel: acc <- acc (+) part, optionally writing the final 16-bit O and L ----
// One thread per 8 output elements of a row; O_acc fp32 [H, n_local, D], part 16-bit.
template <int IS_BF16>
__global__ void __launch_bounds__(256) ring_merge_kernel(
    float *__restrict__ o_acc, const float *__restrict__ l_acc_in, float *__restrict__ l_acc_out,
    const uint16_t *__restrict__ o_part,
    const float *__restrict__ l_part, uint16_t *__restrict__ o_out, float *__restrict__ l_out,
    int n_local, int D, int row_off, int rows, int first, int last) {
  const int vec_per_row = D / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_head = (int64_t)rows * vec_per_row;
  const int h = blockIdx.y;
  if (idx >= per_head) return;
  const int r = row_off + (int)(idx / vec_per_row);
  const int v = (int)(idx % vec_per_row);
  const int64_t row_idx = (int64_t)h * n_local + r;
  const int64_t e = row_idx * D + v * 8;
  const float lp = l_part[row_idx];
  float w_acc = 0.f, w_part = 1.f, l_new = lp;
  if (!first) {
    const float la = l_acc_in[row_idx];
    const float mx = fmaxf(la, lp);
    const float ea = __expf(la - mx), ep = __expf(lp - mx);
    l_new = mx + __logf(ea + ep);
    w_acc = __expf(la - l_new);
    w_part = __expf(lp - l_new);
  }
  const uint4 pv = *reinterpret_cast<const uint4 *>(o_part + e);
  const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
  float out[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float p0, p1;
    if (IS_BF16) {
      p0 = __uint_as_float(pw[i] << 16);
      p1 = __uint_as_float(pw[i] & 0xffff0000u);
    } else {
      asm("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}"
          : "=f"(p0), "=f"(p1) : "r"(pw[i]));
    }
    out[2 * i] = p0 * w_part;
    out[2 * i + 1] = p1 * w_part;
  }
  if (!first) {
    const float4 a0 = *reinterpret_cast<const float4 *>(o_acc + e);
    const float4 a1 = *reinterpret_cast<const float4 *>(o_acc + e + 4);
    out[0] += a0.x * w_acc; out[1] += a0.y * w_acc; out[2] += a0.z * w_acc; out[3] += a0.w * w_acc;
    out[4] += a1.x * w_acc; out[5] += a1.y * w_acc; out[6] += a1.z * w_acc; out[7] += a1.w * w_acc;
  }
  if (!last) {
    *reinterpret_cast<float4 *>(o_acc + e) = make_float4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<float4 *>(o_acc + e + 4) = make_float4(out[4], out[5], out[6], out[7]);
    if (v == 0) l_acc_out[row_idx] = l_new;  // other threads of this row still read l_acc_in
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (IS_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(out[2 * i + 1]), "f"(out[2 * i]));
      else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(out[2 * i + 1]), "f"(out[2 * i]));
    }
    *reinterpret_cast<uint4 *>(o_out + e) = make_uint4(w[0], w[1], w[2], w[3]);
    if (v == 0 && l_out) l_out[row_idx] = l_new;
  }
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

}  // namespace
}  // namespace fa

using namespace fa;

extern "C" {

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

int fa_ring_create(void **ring_out, const void *unique_id, int rank, int world, int device) {
  FA_REQUIRE(ring_out && unique_id && world >= 1 && rank >= 0 && rank < world, "bad ring arguments");
  NcclApi *api = nccl();
  if (!api) return set_error(FA_ERR_NCCL, "libnccl.so.2 could not be loaded");
  FA_CUDA_CHECK(cudaSetDevice(device));
  Ring *r = new Ring();
  r->rank = rank; r->world = world; r->device = device;
  NcclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  int rc = api->CommInitRank(&r->comm, world, id, rank);
  if (rc != 0) {
    delete r;
    return set_error(FA_ERR_NCCL, "ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(rc) : "?");
  }
  FA_CUDA_CHECK(cudaStreamCreateWithFlags(&r->comm_stream, cudaStreamNonBlocking));
  FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->inputs_ready, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->recv_done[i], cudaEventDisableTiming));
    FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->compute_done[i], cudaEventDisableTiming));
    FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->add_done[i], cudaEventDisableTiming));
    FA_CUDA_CHECK(cudaEventCreateWithFlags(&r->acc_recv_done[i], cudaEventDisableTiming));
  }
  *ring_out = r;
  return FA_OK;
}

int fa_ring_destroy(void *ring) {
  Ring *r = reinterpret_cast<Ring *>(ring);
  if (!r) return FA_OK;
  cudaSetDevice(r->device);
  if (r->comm_stream) cudaStreamSynchronize(r->comm_stream);
  if (r->comm && nccl() && nccl()->CommDestroy) nccl()->CommDestroy(r->comm);
  if (r->comm_stream) cudaStreamDestroy(r->comm_stream);
  if (r->inputs_ready) cudaEventDestroy(r->inputs_ready);
  for (int i = 0; i < 2; ++i) {
    if (r->recv_done[i]) cudaEventDestroy(r->recv_done[i]);
    if (r->compute_done[i]) cudaEventDestroy(r->compute_done[i]);
    if (r->add_done[i]) cudaEventDestroy(r->add_done[i]);
    if (r->acc_recv_done[i]) cudaEventDestroy(r->acc_recv_done[i]);
  }
  delete r;
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

size_t fa_ring_workspace_bytes(int n_local, int D, int H, int dtype) {
  (void)dtype;
  if (n_local < 1 || H < 1 || D < 1) return 0;
  const size_t tile = (size_t)H * n_local * D;
  size_t bytes = 2 * (2 * tile * 2)   // two receive slots, each K | V
                 + tile * 2            // partial O (16-bit)
                 + tile * 4            // O accumulator (fp32)
                 + 3 * (size_t)H * n_local * 4;  // partial L, two L accumulators (ping-pong)
  return bytes + 1024;
}

// Workspace of the all-gather forward mode: room for every rank's K and V instead of two receive slots.
size_t fa_ring_workspace_bytes_gather(int world, int n_local, int D, int H, int dtype) {
  const size_t base = fa_ring_workspace_bytes(n_local, D, H, dtype);
  if (base == 0 || world < 1) return 0;
  const size_t tile = (size_t)H * n_local * D;
  return base - 2 * (2 * tile * 2) + 2 * (size_t)world * tile * 2;
}

int fa_ring_attention_forward(void *ring, const void *Q, const void *K, const void *V, void *O, float *L_out,
                              int n_local, int D, int H, float scale, int is_causal, int dtype, void *workspace,
                              size_t workspace_bytes, fa_stream_t stream_) {
  Ring *r = reinterpret_cast<Ring *>(ring);
  FA_REQUIRE(r && Q && K && V && O, "null argument");
  FA_REQUIRE(D == 64 || D == 128, "D must be 64 or 128 (got %d)", D);
  FA_REQUIRE(!is_causal || n_local % 2 == 0, "causal ring attention needs an even n_local");
  FA_REQUIRE(n_local % 8 == 0, "n_local must be a multiple of 8");
  const size_t need = fa_ring_workspace_bytes(n_local, D, H, dtype);
  if (!workspace || workspace_bytes < need)
    return set_error(FA_ERR_WORKSPACE, "ring workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  NcclApi *api = nccl();
  if (!api) return set_error(FA_ERR_NCCL, "libnccl.so.2 could not be loaded");
  cudaStream_t st = (cudaStream_t)stream_;
  const int P = r->world;
  const size_t tile_elems = (size_t)H * n_local * D;
  const size_t tile_bytes = tile_elems * 2;
  char *ws = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  // All-gather mode (opt-in: FA_RING_GATHER=1 and a workspace of fa_ring_workspace_bytes_gather):
  // for steps whose compute is shorter than their K/V hand-off (medium N on many GPUs) the ring is
  // bound by its per-step transfer; here every rank's K/V is gathered once (ncclAllGather on the side
  // stream, under the local block) and the P - 1 remote blocks then run back to back.
  static const int gather_env = [] { const char *e = getenv("FA_RING_GATHER"); return e ? atoi(e) : 0; }();
  const bool gather = gather_env != 0 && P > 1 && api->AllGather != nullptr &&
                      workspace_bytes >= fa_ring_workspace_bytes_gather(P, n_local, D, H, dtype);
  const size_t kv_area = gather ? 2 * (size_t)P * tile_bytes : 4 * tile_bytes;  // [K of all ranks | V of all ranks] or 2 slots
  char *slot[2] = {ws, ws + 2 * tile_bytes};
  char *k_all = ws, *v_all = ws + (size_t)P * tile_bytes;
  uint16_t *o_part = reinterpret_cast<uint16_t *>(ws + kv_area);
  float *o_acc = reinterpret_cast<float *>(ws + kv_area + tile_bytes);
  float *l_part = reinterpret_cast<float *>(ws + kv_area + tile_bytes + tile_elems * 4);
  float *l_acc[2] = {l_part + (size_t)H * n_local, l_part + 2 * (size_t)H * n_local};
  int l_cur[2] = {0, 0};  // which L accumulator holds the current value, per half
  const int next = (r->rank + 1) % P, prev = (r->rank - 1 + P) % P;
  const int64_t hs = (int64_t)n_local * D;

  FA_CUDA_CHECK(cudaEventRecord(r->inputs_ready, st));
  // which local rows have received a contribution so far (for the first/last flags per row range)
  bool touched[2] = {false, false};  // [first half, second half] when causal; [all, -] otherwise
  // ---- local work of ring step s on the chunk (curK, curV) of rank (rank - s) mod P ----
  auto do_step = [&](int s, const void *curK, const void *curV) -> int {
    Block b;
    ring_plan(r->rank, P, s, n_local, is_causal, &b);
    const uint16_t *q = reinterpret_cast<const uint16_t *>(Q) + (int64_t)b.q_off * D;
    const uint16_t *k = reinterpret_cast<const uint16_t *>(curK) + (int64_t)b.k_off * D;
    const uint16_t *v = reinterpret_cast<const uint16_t *>(curV) + (int64_t)b.k_off * D;
    int rc = launch_fwd_tc_rect(q, k, v, o_part + (int64_t)b.q_off * D, l_part + b.q_off, b.q_rows, b.k_rows, D, scale,
                                (int64_t)H * hs, hs, (int64_t)H * hs, hs, b.causal, 1, H, dtype, st);
    if (rc != FA_OK) return rc;
    // merge per half so that "first contribution" / "last contribution" are uniform within a launch
    const int halves = is_causal ? 2 : 1;
    const int hrows = is_causal ? n_local / 2 : n_local;
    for (int hf = 0; hf < halves; ++hf) {
      const int lo = hf * hrows, hi = lo + hrows;
      if (b.q_off >= hi || b.q_off + b.q_rows <= lo) continue;  // this half is not in the block
      // does any later step touch this half?  (causal: first half is only touched while src <= rank)
      bool later = false;
      for (int s2 = s + 1; s2 < P; ++s2) {
        Block b2;
        ring_plan(r->rank, P, s2, n_local, is_causal, &b2);
        if (!(b2.q_off >= hi || b2.q_off + b2.q_rows <= lo)) later = true;
      }
      const int64_t work = (int64_t)hrows * (D / 8);
      dim3 grid((unsigned)((work + 255) / 256), H);
      if (dtype == FA_DTYPE_BF16)
        ring_merge_kernel<1><<<grid, 256, 0, st>>>(o_acc, l_acc[l_cur[hf]], l_acc[l_cur[hf] ^ 1], o_part, l_part, reinterpret_cast<uint16_t *>(O), L_out,
                                                   n_local, D, lo, hrows, !touched[hf], !later);
      else
        ring_merge_kernel<0><<<grid, 256, 0, st>>>(o_acc, l_acc[l_cur[hf]], l_acc[l_cur[hf] ^ 1], o_part, l_part, reinterpret_cast<uint16_t *>(O), L_out,
                                                   n_local, D, lo, hrows, !touched[hf], !later);
      FA_CUDA_CHECK(cudaGetLastError());
      count_launch();
      touched[hf] = true;
      l_cur[hf] ^= 1;
    }
    return FA_OK;
  };

  if (gather) {
    // the previous call's readers of the gather area are ordered before this one by the stream: `st`
    // reached inputs_ready only after them
    FA_CUDA_CHECK(cudaStreamWaitEvent(r->comm_stream, r->inputs_ready, 0));
    FA_NCCL_CHECK(api->GroupStart());
    FA_NCCL_CHECK(api->AllGather(K, k_all, tile_bytes, kNcclUint8, r->comm, r->comm_stream));
    FA_NCCL_CHECK(api->AllGather(V, v_all, tile_bytes, kNcclUint8, r->comm, r->comm_stream));
    FA_NCCL_CHECK(api->GroupEnd());
    FA_CUDA_CHECK(cudaEventRecord(r->recv_done[0], r->comm_stream));
    int rc = do_step(0, K, V);  // the local block runs under the gather
    if (rc != FA_OK) return rc;
    FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->recv_done[0], 0));
    for (int s = 1; s < P; ++s) {
      const int src = ((r->rank - s) % P + P) % P;
      rc = do_step(s, k_all + (size_t)src * tile_bytes, v_all + (size_t)src * tile_bytes);
      if (rc != FA_OK) return rc;
    }
    return FA_OK;
  }

  const void *curK = K, *curV = V;
  for (int s = 0; s < P; ++s) {
    if (s + 1 < P) {
      // ---- ship the chunk we hold to the next rank while we work on it ----
      if (s == 0) FA_CUDA_CHECK(cudaStreamWaitEvent(r->comm_stream, r->inputs_ready, 0));
      // the slot we are about to overwrite was the chunk step s-1 computed on (s >= 2 only)
      if (s >= 2) FA_CUDA_CHECK(cudaStreamWaitEvent(r->comm_stream, r->compute_done[(s - 1) & 1], 0));
      char *dst = slot[s & 1];
      FA_NCCL_CHECK(api->GroupStart());
      FA_NCCL_CHECK(api->Send(curK, tile_bytes, kNcclUint8, next, r->comm, r->comm_stream));
      FA_NCCL_CHECK(api->Send(curV, tile_bytes, kNcclUint8, next, r->comm, r->comm_stream));
      FA_NCCL_CHECK(api->Recv(dst, tile_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
      FA_NCCL_CHECK(api->Recv(dst + tile_bytes, tile_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
      FA_NCCL_CHECK(api->GroupEnd());
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

size_t fa_ring_workspace_bytes_backward(int n_local, int D, int H, int dtype) {
  (void)dtype;
  if (n_local < 1 || H < 1 || D < 1) return 0;
  const size_t tile = (size_t)H * n_local * D;
  size_t bytes = 2 * (2 * tile * 2)        // two K|V receive slots (16-bit)
                 + 2 * tile * 4            // this step's dK|dV block (fp32)
                 + 2 * (2 * tile * 4)      // two outgoing dK|dV accumulators
                 + 2 * tile * 4            // incoming dK|dV accumulator
                 + (size_t)H * n_local * 4;  // delta
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
  NcclApi *api = nccl();
  if (!api) return set_error(FA_ERR_NCCL, "libnccl.so.2 could not be loaded");
  cudaStream_t st = (cudaStream_t)stream_;
  const int P = r->world;
  const size_t tile_elems = (size_t)H * n_local * D;
  const size_t kv_bytes = tile_elems * 2, g_bytes = tile_elems * 4;
  char *ws = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  char *slot[2] = {ws, ws + 2 * kv_bytes};
  float *tmp = reinterpret_cast<float *>(ws + 4 * kv_bytes);                  // [dK | dV] of this step's block
  float *acc_out[2] = {tmp + 2 * tile_elems, tmp + 4 * tile_elems};
  float *acc_in = tmp + 6 * tile_elems;
  float *delta = tmp + 8 * tile_elems;
  const int next = (r->rank + 1) % P, prev = (r->rank - 1 + P) % P;
  const int64_t hs = (int64_t)n_local * D;
  int rc = launch_bwd_delta(O, dO, delta, n_local, D, (int64_t)H * hs, hs, 1, H, dtype, st);
  if (rc != FA_OK) return rc;
  FA_CUDA_CHECK(cudaEventRecord(r->inputs_ready, st));
  const void *curK = K, *curV = V;
  if (P > 1) {  // A(0): our own chunk starts travelling
    FA_CUDA_CHECK(cudaStreamWaitEvent(r->comm_stream, r->inputs_ready, 0));
    FA_NCCL_CHECK(api->GroupStart());
    FA_NCCL_CHECK(api->Send(curK, kv_bytes, kNcclUint8, next, r->comm, r->comm_stream));
    FA_NCCL_CHECK(api->Send(curV, kv_bytes, kNcclUint8, next, r->comm, r->comm_stream));
    FA_NCCL_CHECK(api->Recv(slot[0], kv_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
    FA_NCCL_CHECK(api->Recv(slot[0] + kv_bytes, kv_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
    FA_NCCL_CHECK(api->GroupEnd());
    FA_CUDA_CHECK(cudaEventRecord(r->recv_done[0], r->comm_stream));
  }
  for (int s = 0; s < P; ++s) {
    Block b;
    ring_plan(r->rank, P, s, n_local, is_causal, &b);
    const uint16_t *q = reinterpret_cast<const uint16_t *>(Q) + (int64_t)b.q_off * D;
    const uint16_t *g = reinterpret_cast<const uint16_t *>(dO) + (int64_t)b.q_off * D;
    const uint16_t *k = reinterpret_cast<const uint16_t *>(curK) + (int64_t)b.k_off * D;
    const uint16_t *v = reinterpret_cast<const uint16_t *>(curV) + (int64_t)b.k_off * D;
    float *blk_dk = (P == 1 ? dK : tmp) + (int64_t)b.k_off * D;
    float *blk_dv = (P == 1 ? dV : tmp + tile_elems) + (int64_t)b.k_off * D;
    rc = launch_bwd_tc_rect(q, k, v, g, L + b.q_off, delta + b.q_off, dQ + (int64_t)b.q_off * D, blk_dk, blk_dv, b.q_rows,
                            b.k_rows, D, scale, (int64_t)H * hs, hs, (int64_t)H * hs, hs, b.causal, s > 0, 1, H, dtype, st);
    if (rc != FA_OK) return rc;
    if (P == 1) break;
    // ---- dK/dV accumulator of the chunk we hold: add our block, pass it on ----
    if (s > 0) FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->acc_recv_done[s & 1], 0));
    {
      const int64_t vecs = (int64_t)tile_elems / 4;
      dim3 grid((unsigned)((vecs + 255) / 256), 2);
      ring_dkv_add_kernel<<<grid, 256, 0, st>>>(acc_out[s & 1], acc_in, tmp, (int64_t)tile_elems, n_local, D, b.k_off,
                                                b.k_rows, s > 0);
      FA_CUDA_CHECK(cudaGetLastError());
      count_launch();
    }
    FA_CUDA_CHECK(cudaEventRecord(r->add_done[s & 1], st));
    FA_CUDA_CHECK(cudaStreamWaitEvent(r->comm_stream, r->add_done[s & 1], 0));
    FA_NCCL_CHECK(api->GroupStart());
    FA_NCCL_CHECK(api->Send(acc_out[s & 1], g_bytes, kNcclUint8, next, r->comm, r->comm_stream));
    FA_NCCL_CHECK(api->Send(acc_out[s & 1] + tile_elems, g_bytes, kNcclUint8, next, r->comm, r->comm_stream));
    if (s + 1 < P) {
      FA_NCCL_CHECK(api->Recv(acc_in, g_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
      FA_NCCL_CHECK(api->Recv(acc_in + tile_elems, g_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
    } else {  // last hop: the finished gradients of our own chunk come home
      FA_NCCL_CHECK(api->Recv(dK, g_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
      FA_NCCL_CHECK(api->Recv(dV, g_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
    }
    FA_NCCL_CHECK(api->GroupEnd());
    FA_CUDA_CHECK(cudaEventRecord(r->acc_recv_done[(s + 1) & 1], r->comm_stream));
    // ---- K/V: switch to the chunk that arrived, forward it if somebody still needs it ----
    if (s + 1 < P) {
      FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->recv_done[s & 1], 0));
      curK = slot[s & 1];
      curV = slot[s & 1] + kv_bytes;
      if (s + 2 < P) {  // the other slot was read by step s, which the comm stream has already waited for
        char *dst = slot[(s + 1) & 1];
        FA_NCCL_CHECK(api->GroupStart());
        FA_NCCL_CHECK(api->Send(curK, kv_bytes, kNcclUint8, next, r->comm, r->comm_stream));
        FA_NCCL_CHECK(api->Send(curV, kv_bytes, kNcclUint8, next, r->comm, r->comm_stream));
        FA_NCCL_CHECK(api->Recv(dst, kv_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
        FA_NCCL_CHECK(api->Recv(dst + kv_bytes, kv_bytes, kNcclUint8, prev, r->comm, r->comm_stream));
        FA_NCCL_CHECK(api->GroupEnd());
        FA_CUDA_CHECK(cudaEventRecord(r->recv_done[(s + 1) & 1], r->comm_stream));
      }
    }
  }
  if (P > 1) FA_CUDA_CHECK(cudaStreamWaitEvent(st, r->acc_recv_done[P & 1], 0));
  return FA_OK;
}

}  // extern "C"
