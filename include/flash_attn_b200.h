/*
 * flash_attn_b200.h -- C ABI of libflash_attn_b200.so (NVIDIA B200, sm_100a).
 *
 * The reference (2thleZ/flash_attention_metal) has no library API: its operator
 * surface is the Metal dispatch convention of main.mm -- a kernel is chosen by
 * its *string name*, its arguments are bound by *buffer index*, and the caller
 * picks the grid.  Every entry point below replaces one such dispatch site and
 * keeps the kernel's name and its argument order (buffer index order).  Paths
 * are relative to the reference root.
 *
 * Common conventions:
 *   - tensors are dense row-major [N, D] per head, D contiguous; batched calls
 *     address head (b, h) at element offset b*batch_stride + h*head_stride
 *     (kernels.metal:622, 932); L is contiguous [B, H, N] (kernels.metal:623).
 *   - `scale` multiplies the dot product (kernels.metal:40, 146, 563, 763).
 *   - causal: key j is excluded for query i when j > i (kernels.metal:748).
 *   - all data pointers are DEVICE pointers, caller-owned, 16-byte aligned.  The
 *     single-GPU entry points allocate nothing (scratch comes from the caller's
 *     workspace) and keep no state between calls beyond a per-thread cache of
 *     encoded TMA descriptors and the process-wide backward-algorithm setting;
 *     the objects that own device memory say so (fa_ring_t: its peer window,
 *     fa_mgpu_t: streams, scratch and windows, fa_host_*: a scratch pool per device).
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL =
 *     the default stream) and return 0 on success or a negative FA_ERR_* code;
 *     fa_last_error() gives the message (thread-local).  The reference exits the
 *     process on error (main.mm:16-22); a library must not.
 *   - differences from the reference, all additive: `is_causal` on the fp32
 *     variants (the reference's fp32 kernels have none, kernels.metal:13-20);
 *     explicit B and H (Metal's grid carried them, main.mm:848, 1001); 64-bit
 *     strides (kernels.metal:608-609 uses int, which overflows at N = 1M);
 *     D in {64, 128} (reference: 64 only, main.mm:12); dtype = fp16 or bf16
 *     (reference: fp16 only).
 *   - results are deterministic: the same call (same shapes, same device) returns
 *     the same bits every time, forward and backward (the reference's backward
 *     is not: float atomics, kernels.metal:1227, 1243).  They are not
 *     bit-identical ACROSS call shapes: a 16-bit forward launch too small to
 *     fill the GPU splits the keys of a row block over a thread-block cluster
 *     and merges the parts, which sums in a different order than one CTA does.
 *
 * This header is plain C: no CUDA types appear in it.
 */
#ifndef FLASH_ATTN_B200_H_
#define FLASH_ATTN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *fa_stream_t; /* cudaStream_t */

enum { FA_DTYPE_FP16 = 0, FA_DTYPE_BF16 = 1 };

enum {
  FA_OK = 0,
  FA_ERR_INVALID = -1,    /* bad argument (shape, alignment, dtype, null pointer) */
  FA_ERR_CUDA = -2,       /* a CUDA runtime/driver call failed */
  FA_ERR_WORKSPACE = -3,  /* workspace missing or too small */
  FA_ERR_UNSUPPORTED = -4,/* valid request the library does not implement */
  FA_ERR_NCCL = -5        /* an NCCL call failed */
};

/* ---- fp32 variants: single head, Q/K/V/O float [N, D] ----------------------
 * Buffer indices 0..6 of the reference = Q, K, V, O, N, D, scale.            */

/* replaces naive_attention_kernel, kernels.metal:12-64, dispatched at
 * main.mm:162-191 and 678-694.  One thread per query row, two passes over the
 * keys, no shared memory: kept deliberately naive, it is the denominator of
 * the speed-up columns in benchmark_results.csv. */
int naive_attention(const float *Q, const float *K, const float *V, float *O, int N, int D,
                    float scale, int is_causal, fa_stream_t stream);

/* replaces flash_attention_kernel (V1, tiled), kernels.metal:72-171, dispatched
 * at main.mm:198-224 and 708-724. */
int flash_attention(const float *Q, const float *K, const float *V, float *O, int N, int D,
                    float scale, int is_causal, fa_stream_t stream);

/* replaces flash_attention_v2_kernel (V2, float4 + double buffering),
 * kernels.metal:462-596, dispatched at main.mm:259-275 and 738-754. */
int flash_attention_v2(const float *Q, const float *K, const float *V, float *O, int N, int D,
                       float scale, int is_causal, fa_stream_t stream);

/* Batched form of flash_attention_v2 (the
 * reference's "high occupancy" table has a FlashV2 column it never fills,
 * main.mm:883, 1202).  Heads addressed as in the 16-bit kernels. */
int flash_attention_v2_batched(const float *Q, const float *K, const float *V, float *O, int N,
                               int D, float scale, int64_t batch_stride, int64_t head_stride,
                               int is_causal, int B, int H, fa_stream_t stream);

/* ---- 16-bit variants (tensor cores: tcgen05 + TMEM + TMA) ------------------ */

/* replaces flash_attention_simd_kernel (V3), kernels.metal:177-455, dispatched
 * at main.mm:332-349 and 777-796: single head, no L, no causal mask. */
int flash_attention_simd(const void *Q, const void *K, const void *V, void *O, int N, int D,
                         float scale, int dtype, fa_stream_t stream);

/* replaces flash_attention_v4_half_kernel (V4), kernels.metal:600-883,
 * dispatched at main.mm:414-440, 520-546, 819-855, 977-1008.  Buffer indices
 * 0..10 = Q, K, V, O, N, D, scale, batch_stride, head_stride, L_out, is_causal.
 * L_out[b, h, i] = max_j(s_ij) + log(sum_j exp(s_ij - max)) over the scaled
 * scores, natural log (kernels.metal:863); may be NULL. */
int flash_attention_v4_half(const void *Q, const void *K, const void *V, void *O, int N, int D,
                            float scale, int64_t batch_stride, int64_t head_stride, float *L_out,
                            int is_causal, int B, int H, int dtype, fa_stream_t stream);

/* replaces flash_attention_backward_kernel, kernels.metal:905-1265, dispatched
 * at main.mm:1027-1061.  Buffer indices 0..14 = Q, K, V, O, dO (16-bit), L
 * (float), dQ, dK, dV (float), N, D, scale, batch_stride, head_stride,
 * is_causal.  All three gradients are fully overwritten (the reference needs
 * dK/dV pre-zeroed because it accumulates them with float atomics,
 * kernels.metal:1227, 1243; this library uses none and is run-to-run
 * deterministic).  `workspace` holds D_i = sum_d O*dO (kernels.metal:983-990);
 * size it with fa_workspace_bytes_backward(). */
int flash_attention_backward(const void *Q, const void *K, const void *V, const void *O,
                             const void *dO, const float *L, float *dQ, float *dK, float *dV,
                             int N, int D, float scale, int64_t batch_stride,
                             int64_t head_stride, int is_causal, int B, int H, int dtype,
                             void *workspace, size_t workspace_bytes, fa_stream_t stream);

size_t fa_workspace_bytes_backward(int N, int D, int B, int H);

/* Two implementations of the backward, same arguments, same tolerances, both free of float atomics and
 * bitwise reproducible (process-wide choice; the workspace size covers both):
 *   FA_BWD_TWO_KERNEL (default)  one kernel owns every dK/dV tile, another every dQ tile; S and dP are
 *                                recomputed in both (7 GEMM-units for the 5 of the algorithm).
 *   FA_BWD_FUSED                 one kernel, 5 GEMMs per tile pair; the dQ contributions of the key tiles
 *                                are added by TMA add-reductions in a fixed (ascending key tile) order
 *                                enforced with counters in the workspace.  One launch instead of two, less
 *                                tensor work, but bound by the ordered L2 reduction: measured slower on
 *                                B200 except for very small problems. */
enum { FA_BWD_TWO_KERNEL = 0, FA_BWD_FUSED = 1 };
int fa_set_backward_algorithm(int algorithm);
int fa_get_backward_algorithm(void);

/* Rectangular (cross-attention) form of flash_attention_v4_half: Nq query rows against Nk
 * keys/values, non-causal, separate strides for Q/O and K/V (the
 * reference only has Nq == Nk).  Ring attention is built on it. */
int flash_attention_v4_half_rect(const void *Q, const void *K, const void *V, void *O, int Nq, int Nk,
                                 int D, float scale, int64_t q_batch_stride, int64_t q_head_stride,
                                 int64_t kv_batch_stride, int64_t kv_head_stride, float *L_out, int B,
                                 int H, int dtype, fa_stream_t stream);

/* Rectangular backward (cross-attention gradients; SURVEY.md section 8 row f3, the reference's tail
 * handling at kernels.metal:103-110, 658-662 is the only hook it has): gradients of
 * attention(Q[Nq], K[Nk], V[Nk]), non-causal.  L is the log-sum-exp written by the forward over the
 * FULL key range of each query row, `delta` = rowsum(O o dO) over the full output (fa_rowsum_delta);
 * so the gradients of a softmax whose keys are processed in several chunks are the SUM of one call per
 * chunk (dQ with acc_dq = 1 after the first; dK/dV per chunk).  dK/dV may both be NULL (dQ only) and dQ
 * may be NULL (dK/dV only). */
int flash_attention_backward_rect(const void *Q, const void *K, const void *V, const void *dO, const float *L,
                                  const float *delta, float *dQ, float *dK, float *dV, int Nq, int Nk, int D,
                                  float scale, int64_t q_batch_stride, int64_t q_head_stride,
                                  int64_t kv_batch_stride, int64_t kv_head_stride, int acc_dq, int B, int H,
                                  int dtype, fa_stream_t stream);
/* delta[b, h, i] = sum_d O[b,h,i,d] * dO[b,h,i,d] (kernels.metal:983-990), addressed like L */
int fa_rowsum_delta(const void *O, const void *dO, float *delta, int N, int D, int64_t batch_stride,
                    int64_t head_stride, int B, int H, int dtype, fa_stream_t stream);

/* ---- ring / context-parallel attention across the GPUs of one box (new; not in the
 * reference, BASELINE.json config 5).  One process per GPU (for one process driving all GPUs see
 * fa_mgpu_* below).  Each rank holds n_local rows of Q, K, V per head, contiguous [H, n_local, D].
 * Non-causal: rank r holds global rows [r*n_local, (r+1)*n_local).  Causal: zig-zag -- with
 * c = n_local/2, local rows [0,c) are global chunk r and local rows [c,2c) are global chunk
 * 2*world-1-r (fa_ring_local_rows), so every rank does the same amount of unmasked work at every
 * step.  The partial result of a step is folded into the running (O, L) inside the forward kernel's
 * epilogue: one launch per step.
 *
 * Transport -- how K/V chunks (and, backward, dK/dV accumulators) move between GPUs.  It is a property
 * of the ring, fixed at creation and agreed by all ranks (every rank must pass the same value):
 *   PEER        each rank exposes a window of device memory to its peers (CUDA IPC); chunks are
 *               pulled straight from their owner by the copy engines over NVLink and ranks
 *               synchronise through 32-bit flags with stream memory operations.  Uses no SMs.
 *   NCCL        ncclSend/ncclRecv ring on a side stream.
 *   NCCL_GATHER forward only: one ncclAllGather of every rank's K/V, then the remote blocks back to
 *               back (needs the larger workspace of fa_ring_workspace_bytes_ex).
 *   AUTO        PEER when every rank can set it up, else NCCL. */
typedef void *fa_ring_t;
enum {
  FA_RING_TRANSPORT_AUTO = 0,
  FA_RING_TRANSPORT_NCCL = 1,
  FA_RING_TRANSPORT_NCCL_GATHER = 2,
  FA_RING_TRANSPORT_PEER = 3
};
int fa_ring_unique_id_bytes(void);
int fa_ring_get_unique_id(void *out, int bytes);            /* rank 0; ship the bytes to the others */
int fa_ring_create(fa_ring_t *ring, const void *unique_id, int rank, int world, int device); /* AUTO */
int fa_ring_create_ex(fa_ring_t *ring, const void *unique_id, int rank, int world, int device, int transport);
int fa_ring_transport(fa_ring_t ring);                      /* the transport in effect (never AUTO) */
int fa_ring_destroy(fa_ring_t ring);
/* forward workspace for a ring of `world` ranks using `transport` (as returned by fa_ring_transport) */
size_t fa_ring_workspace_bytes_ex(int world, int transport, int n_local, int D, int H, int dtype);
size_t fa_ring_workspace_bytes(int n_local, int D, int H, int dtype);             /* PEER and NCCL */
size_t fa_ring_workspace_bytes_gather(int world, int n_local, int D, int H, int dtype); /* NCCL_GATHER */
/* A workspace smaller than the ring's transport needs is an error (FA_ERR_WORKSPACE), never a silent
 * change of algorithm.  Collective: every rank calls with the same shapes. */
int fa_ring_attention_forward(fa_ring_t ring, const void *Q, const void *K, const void *V, void *O,
                              float *L_out, int n_local, int D, int H, float scale, int is_causal,
                              int dtype, void *workspace, size_t workspace_bytes, fa_stream_t stream);
/* Backward of the same layout: L is the log-sum-exp written by fa_ring_attention_forward, O its
 * output.  dQ, dK, dV are fp32 [H, n_local, D], fully overwritten, deterministic.  K/V chunks
 * travel one step ahead of the tile loop, each chunk's dK/dV accumulator one step behind it. */
size_t fa_ring_workspace_bytes_backward(int n_local, int D, int H, int dtype);
int fa_ring_attention_backward(fa_ring_t ring, const void *Q, const void *K, const void *V, const void *O,
                               const void *dO, const float *L, float *dQ, float *dK, float *dV,
                               int n_local, int D, int H, float scale, int is_causal, int dtype,
                               void *workspace, size_t workspace_bytes, fa_stream_t stream);
/* host-only helpers (no GPU needed): the block a rank computes at a ring step, in local row
 * coordinates, and the global rows a rank owns */
int fa_ring_plan(int rank, int world, int step, int n_local, int is_causal, int *src_rank, int *q_off,
                 int *q_rows, int *k_off, int *k_rows, int *block_causal);
int fa_ring_local_rows(int rank, int world, int n_local, int is_causal, int64_t first_row[2], int rows[2]);
/* how the forward launch of a step folds its partial result into the running (O, L): merge mode (0 the only
 * partial, 1 first, 2 middle, 3 last) of the launch's rows below / from half_rows */
int fa_ring_merge_plan(int rank, int world, int step, int n_local, int is_causal, int *lo, int *hi, int *half_rows);

/* ---- fa_mgpu_*: ONE process driving several GPUs of one box (the reference's
 * harness is a single process, main.mm:881-1204).  The group owns a stream per device, its scratch
 * memory and, for ring attention, the peer windows (plain peer access: no NCCL, no IPC).  Pointer
 * arguments are arrays with one DEVICE pointer per group device, each on its own device.  Calls only
 * enqueue work on the group's streams; fa_mgpu_synchronize waits for all devices. */
typedef void *fa_mgpu_t;
int fa_mgpu_create(fa_mgpu_t *group, const int *devices, int n_devices);
int fa_mgpu_destroy(fa_mgpu_t group);
int fa_mgpu_device_count(fa_mgpu_t group);
fa_stream_t fa_mgpu_stream(fa_mgpu_t group, int index);   /* stream of the index-th device of the group */
int fa_mgpu_synchronize(fa_mgpu_t group);
/* batch x heads sharding (BASELINE config 4): device i runs heads[i] independent heads, its tensors are
 * contiguous [heads[i], N, D]; no communication (kernels.metal:622). */
int fa_mgpu_sharded_forward(fa_mgpu_t group, const void *const *Q, const void *const *K, const void *const *V,
                            void *const *O, float *const *L, int N, int D, float scale, int is_causal,
                            const int *heads, int dtype);
int fa_mgpu_sharded_backward(fa_mgpu_t group, const void *const *Q, const void *const *K, const void *const *V,
                             const void *const *O, const void *const *dO, const float *const *L, float *const *dQ,
                             float *const *dK, float *const *dV, int N, int D, float scale, int is_causal,
                             const int *heads, int dtype);
/* ring / context-parallel attention (BASELINE config 5), layout as fa_ring_attention_*: device i is rank i */
int fa_mgpu_ring_forward(fa_mgpu_t group, const void *const *Q, const void *const *K, const void *const *V,
                         void *const *O, float *const *L, int n_local, int D, int H, float scale, int is_causal,
                         int dtype);
int fa_mgpu_ring_backward(fa_mgpu_t group, const void *const *Q, const void *const *K, const void *const *V,
                          const void *const *O, const void *const *dO, const float *const *L, float *const *dQ,
                          float *const *dK, float *const *dV, int n_local, int D, int H, float scale,
                          int is_causal, int dtype);

/* ---- host-buffer entry points ----------------------------------------------
 * The reference's buffers are MTLResourceStorageModeShared (main.mm:104-115):
 * the host writes inputs and reads outputs in place.  These calls give a
 * caller with HOST tensors the same one-call behaviour: copy in, run, copy out,
 * synchronise.  Contiguous [B, H, N, D].  Device scratch comes from a grow-only
 * pool inside the library (fa_host_release frees it); the 16-bit calls pipeline
 * copies and kernels over groups of heads.  Pinned host memory gives full PCIe
 * speed.  These are what the harness-style caller and bench.py's end-to-end leg use. */
int fa_host_attention_f32(int variant /*0 naive, 1 v1, 2 v2*/, const float *Q, const float *K,
                          const float *V, float *O, int N, int D, float scale, int is_causal);
int fa_host_attention_half(const void *Q, const void *K, const void *V, void *O, float *L_out,
                           int N, int D, float scale, int is_causal, int B, int H, int dtype);
/* forward (O, L_out) followed by backward (dQ, dK, dV in fp32) in one call */
int fa_host_attention_fwd_bwd_half(const void *Q, const void *K, const void *V, const void *dO,
                                   void *O, float *L_out, float *dQ, float *dK, float *dV, int N,
                                   int D, float scale, int is_causal, int B, int H, int dtype);
/* The same with a choice of gradient type on the way out: grad_dtype = -1 keeps the reference's fp32
 * gradients (kernels.metal:912-914 `atomic_float* dQ/dK/dV`), FA_DTYPE_FP16 / FA_DTYPE_BF16 rounds them on
 * the device (round to nearest even) so half as many bytes cross PCIe -- the device->host copy of the
 * fp32 gradients is what bounds the fp32 form of this call.  dQ/dK/dV then point to 16-bit buffers. */
int fa_host_attention_fwd_bwd_half_ex(const void *Q, const void *K, const void *V, const void *dO, void *O,
                                      float *L_out, void *dQ, void *dK, void *dV, int N, int D, float scale,
                                      int is_causal, int B, int H, int dtype, int grad_dtype);
void fa_host_release(void);

/* ---- support ----------------------------------------------------------------*/
const char *fa_last_error(void);
int fa_version(void);          /* major*10000 + minor*100 + patch */
int fa_device_count(void);     /* CUDA devices visible; <0 on error */
/* Loads every 16-bit and ring kernel into the current device's context ahead of time (CUDA loads a kernel
 * lazily at its first launch, which synchronises the context).  Optional; fa_mgpu_create does it. */
int fa_preload_kernels(void);
/* Kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
long fa_launch_count(void);
void fa_reset_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FLASH_ATTN_B200_H_ */
