// 16-bit FlashAttention backward for sm_100a (tcgen05 + TMEM + TMA), deterministic:
// no float atomics anywhere (the reference accumulates dK/dV with them,
// kernels.metal:1227, 1243).
//
// Replaces flash_attention_backward_kernel (kernels.metal:905-1265).  Same math:
//   D_i  = sum_d O_id dO_id                         (kernels.metal:983-990)
//   P    = exp(scale * Q K^T - L_i)                 (kernels.metal:1082-1089)
//   dV  += P^T dO        dP = dO V^T
//   dS   = P o (dP - D_i) * scale                   (kernels.metal:1160-1169)
//   dQ  += dS K          dK += dS^T Q
// split into three launches so that every gradient tile has exactly one owner CTA:
//   1. bwd_delta_kernel : D_i into the caller's workspace (warp-shuffle row sums)
//   2. bwd_dkdv_kernel  : one CTA per 128-key tile, streams 128-row Q/dO tiles,
//                         S^T = K Q^T and dP^T = V dO^T land transposed in TMEM so that
//                         P^T and dS^T are directly the TMEM A operands of
//                         dV += P^T dO and dK += dS^T Q  (4 GEMMs per tile pair)
//   3. bwd_dq_kernel    : one CTA per 2 x 128 query rows, streams 128-row K/V tiles,
//                         S = Q K^T, dP = dO V^T, dQ += dS K  (3 GEMMs per tile pair)
// Every MMA is M = 128, N = 128 (or N = D): tools/mma_rate_probe.cu shows N = 64
// shared-memory-operand MMAs run at 67 % because the A tile is re-read per 64 columns.
// Both kernels split the element-wise work in two phases (P from S, then dS from dP) so the
// tensor core computes dP while the exponentials run, and dV / the next S while dS is formed.
// S and dP are recomputed in both kernels (7 GEMMs instead of 5): that is the price of
// an atomic-free, order-independent dQ.
//
// Warp roles in both kernels: warps 0-3 and 4-7 are two element-wise warpgroups (one
// thread per TMEM lane), warp 8 issues MMAs, warp 9 drives TMA, warps 10-11 only complete the
// warpgroup so setmaxnreg can move registers to the element-wise warps.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fa_internal.h"
#include "sched.cuh"
#include "sm100_ptx.cuh"
#include "tensormap.h"

namespace fa {
namespace {

using namespace ptx;

constexpr int kBwdThreads = 384;
constexpr int kMmaWarp = 8;
constexpr int kLoadWarp = 9;
constexpr float kLog2e = 1.4426950408889634f;
// Every FA_BWD_EMU-th pair of exponentials runs on the FMA pipe instead of MUFU.EX2 (0 = none), as in
// the forward kernel; P^T / dS^T / dS are handed to the MMA warp in parts so the accumulating MMAs
// start while the rest of the tile is still being computed.
#ifndef FA_BWD_EMU
#define FA_BWD_EMU 3
#endif
#ifndef FA_BWD_EMU_DQ
#define FA_BWD_EMU_DQ FA_BWD_EMU
#endif
#ifndef FA_BWD_SPLIT
#define FA_BWD_SPLIT 1
#endif
constexpr int kBwdEmu = FA_BWD_EMU;
constexpr int kDkdvParts = FA_BWD_SPLIT ? 2 : 1;  // hand-offs per phase in bwd_dkdv_kernel
constexpr int kDqParts = FA_BWD_SPLIT ? 4 : 1;    // dS hand-offs per item in bwd_dq_kernel
template <int kEvery = kBwdEmu>
__device__ __forceinline__ bool emulate_pair(int pair_index) {
  return kEvery > 0 && (pair_index % (kEvery > 0 ? kEvery : 1)) == (kEvery > 0 ? kEvery : 1) - 1;
}
// 2^x of a packed pair: MUFU.EX2 or the FMA-pipe polynomial
__device__ __forceinline__ uint64_t exp2_pair(uint64_t x2, bool emulate) {
  return emulate ? exp2_emulated_x2(x2) : pack_f32x2(ex2(lo_f32(x2)), ex2(hi_f32(x2)));
}

// FA_BWD_TRACE (development builds only): the first dQ CTA records clock64() timestamps of its
// hand-offs for items s = 8..11 of both tiles into the buffer set with fa_debug_set_prof_buffer.
#ifdef FA_BWD_TRACE
#define FA_BTRACE(cond, s_, slot)                                                                \
  do {                                                                                           \
    if (p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (cond) &&  \
        (s_) >= 8 && (s_) < 12)                                                                  \
      p.prof[64 + ((s_) - 8) * 64 + (slot)] = clock64();                                         \
  } while (0)
#else
#define FA_BTRACE(cond, s_, slot) do {} while (0)
#endif

struct BwdParams {
  const float *L;      // [B, H, N] log-sum-exp of the scaled scores (natural log)
  const float *delta;  // [B, H, N] D_i (workspace)
  float *dQ, *dK, *dV;
  int Nq, Nk, H;      // query rows / keys (equal except for ring-attention blocks)
  float scale, scale_log2;
  int64_t batch_stride, head_stride;        // elements, of Q / O / dO / dQ (and L, delta via / D)
  int64_t kv_batch_stride, kv_head_stride;  // elements, of K / V / dK / dV
  int causal;         // requires Nq == Nk
  int acc_dq;         // dQ += instead of dQ = (ring attention accumulates over K/V chunks)
  int group, n_heads; // dispatch order (sched.cuh)
#ifdef FA_BWD_TRACE
  long long *prof;  // phase-timing buffer (trace builds only)
#endif
};

// ---------------------------------------------------------------------------
// 1. D_i = sum_d O_id * dO_id : one warp per row, 128-bit loads, shuffle reduce
// ---------------------------------------------------------------------------
template <int D, int IS_BF16>
__global__ void __launch_bounds__(256) bwd_delta_kernel(const uint16_t *__restrict__ O,
                                                         const uint16_t *__restrict__ dO,
                                                         float *__restrict__ delta, int N, int H,
                                                         int64_t batch_stride, int64_t head_stride) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= N) return;
  const int64_t off = (int64_t)blockIdx.z * batch_stride + (int64_t)blockIdx.y * head_stride;
  const uint16_t *o = O + off + (int64_t)row * D;
  const uint16_t *g = dO + off + (int64_t)row * D;
  float acc = 0.f;
  constexpr int kVecs = D / 8;  // uint4 = 8 elements
  for (int v = lane; v < kVecs; v += 32) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(o) + v);
    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(g) + v);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a0, a1, b0, b1;
      if (IS_BF16) {
        a0 = __uint_as_float(aw[i] << 16); a1 = __uint_as_float(aw[i] & 0xffff0000u);
        b0 = __uint_as_float(bw[i] << 16); b1 = __uint_as_float(bw[i] & 0xffff0000u);
      } else {
        const __half2 ha = *reinterpret_cast<const __half2 *>(&aw[i]);
        const __half2 hb = *reinterpret_cast<const __half2 *>(&bw[i]);
        a0 = __low2float(ha); a1 = __high2float(ha); b0 = __low2float(hb); b1 = __high2float(hb);
      }
      acc = fmaf(a0, b0, acc);
      acc = fmaf(a1, b1, acc);
    }
  }
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, w);
  if (lane == 0) delta[off / D + row] = acc;
}

// Shared-memory geometry.  Every tile is [128 rows][D] stored as D/64 chunks of
// [128 rows][64 elements] (128-byte rows, 128-byte swizzle) exactly as TMA writes them.
template <int D>
struct BwdCfg {
  static constexpr int kChunks = D / 64;
  static constexpr int kChunk = 128 * 128;           // bytes of one chunk
  static constexpr int kTile = kChunks * kChunk;     // one 128-row tile
};

// K-major operand, k-step kk (16 elements of the head dim): chunk kk/4, 32 bytes per step
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t tile_addr, int kk) {
  return make_sdesc_sw128(tile_addr + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
}
// MN-major B operand over a [128 rows][D] tile: k-step kk covers rows 16kk..16kk+15
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t tile_addr, int kk) {
  return make_sdesc_sw128(tile_addr + kk * 2048, 128 * 128, 1024);
}

template <int D>
__device__ __forceinline__ void tma_load_tile(unsigned char *dst, const CUtensorMap *map, uint64_t *bar, int row, int h,
                                              int b) {
#pragma unroll
  for (int c = 0; c < D / 64; ++c) tma_load_4d(dst + c * (128 * 128), map, bar, c * 64, row, h, b);
}

// ---------------------------------------------------------------------------
// 2. dK / dV : CTA owns keys [128 j, 128 j + 128); streams 128-row Q_i / dO_i tiles.
//    TMEM: X = S^T [0,128)   Y = dP^T [128,256)   dV [256,256+D)   dK [256+D,256+2D)
//    Warpgroup w owns query columns [64w, 64w+64) of X and Y; its 16-bit P^T / dS^T go back
//    over the first 32 columns of its own half, so k-steps 0-3 read columns [0,32) and k-steps
//    4-7 read columns [64,96) of X (resp. Y).
//    MMA order:  X(0) Y(0) | P? dV(0) X(1) | dS? dK(0) Y(1) | P? dV(1) X(2) | ...
// ---------------------------------------------------------------------------
template <int D>
struct DkdvCfg : BwdCfg<D> {
  static constexpr int kQSlots = D == 128 ? 3 : 4;
  static constexpr int kDoSlots = D == 128 ? 2 : 4;
  static constexpr int kSmemTiles = (2 + kQSlots + kDoSlots) * BwdCfg<D>::kTile;
  static constexpr int kStatBytes = 2 * 2 * 128 * 4;  // [wg][double buffer][L*log2e 64 | D*scale 64]
  // 224 KB of tiles at D = 128: only 512 bytes of alignment slack fit under the 227 KB limit (the
  // dynamic shared window starts 1 KB-aligned in practice; the kernel traps if it ever does not)
  static constexpr int kSmemBytes = kSmemTiles + kStatBytes + 512 + 256;
};

template <int D, int IS_BF16>
__global__ void __launch_bounds__(kBwdThreads, 1)
bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const BwdParams p) {
  using Cfg = DkdvCfg<D>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  if (smem - smem_raw > 512) __trap();
  unsigned char *sK = smem;
  unsigned char *sV = smem + Cfg::kTile;
  unsigned char *sQ = smem + 2 * Cfg::kTile;                      // [kQSlots]
  unsigned char *sDO = sQ + Cfg::kQSlots * Cfg::kTile;            // [kDoSlots]
  float *sLD = reinterpret_cast<float *>(smem + Cfg::kSmemTiles);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemTiles + Cfg::kStatBytes);
  uint64_t *res_full = bars;        // K_j, V_j landed
  uint64_t *acc_full = bars + 1;    // dV, dK final
  uint64_t *x_full = bars + 2;      // S^T ready
  uint64_t *y_full = bars + 3;      // dP^T ready
  uint64_t *p_ready = bars + 4;     // [2] P^T part stored by both warpgroups
  uint64_t *ds_ready = bars + 6;    // [2] dS^T part stored by both warpgroups
  uint64_t *q_full = bars + 8;                       // [kQSlots]
  uint64_t *q_empty = q_full + Cfg::kQSlots;
  uint64_t *do_full = q_empty + Cfg::kQSlots;        // [kDoSlots]
  uint64_t *do_empty = do_full + Cfg::kDoSlots;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(do_empty + Cfg::kDoSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // dispatch order: see sched.cuh (causal: key tile 0 sees every query tile and goes first)
  const BlockCoord bc = decode_block(p.group, p.n_heads, p.H);
  if (bc.b < 0) return;
  const int j = bc.blk, h = bc.h, b = bc.b;
  const int key0 = j * 128;
  const int n_tiles_all = (p.Nq + 127) / 128;
  const int i_start = p.causal ? j : 0;    // first query tile that sees these keys
  const int n = n_tiles_all - i_start;     // >= 1 because key0 < Nk (and Nq == Nk when causal)
  const int64_t kv_off = (int64_t)b * p.kv_batch_stride + (int64_t)h * p.kv_head_stride;
  const int64_t vec_off = ((int64_t)b * p.batch_stride + (int64_t)h * p.head_stride) / D;

  if (threadIdx.x == 0) {
    mbar_init(res_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(x_full, 1);
    mbar_init(y_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&p_ready[i], 8 * kArrivalsPerWarp); mbar_init(&ds_ready[i], 8 * kArrivalsPerWarp); }
    for (int i = 0; i < Cfg::kQSlots; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < Cfg::kDoSlots; ++i) { mbar_init(&do_full[i], 1); mbar_init(&do_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ======================= element-wise warpgroups =======================
    setmaxnreg_inc<216>();
    const int wg = warp >> 2;
    const int tid = (warp & 3) * 32 + lane;  // TMEM lane = key row within the tile
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tX = tmem_base + lane_off + wg * 64;         // this warpgroup's half of S^T / P^T
    const uint32_t tY = tmem_base + lane_off + 128 + wg * 64;   // ... of dP^T / dS^T
    const int key = key0 + tid;
    // per-column statistics (-L_i * log2e for threads 0-63, -D_i * scale for threads 64-127) of the 64
    // query columns this warpgroup owns, fetched one tile ahead
    // The load is issued one tile ahead and its value is only touched (scaled) at the top of the
    // next iteration: any arithmetic on it here would stall this in-order warp for the full global
    // load latency every tile (measured: ~500 of 2700 cycles per tile).
    const float *stat_src = (tid < 64 ? p.L : p.delta) + vec_off + wg * 64 + (tid & 63);
    const float stat_coef = tid < 64 ? -kLog2e : -p.scale;
    auto fetch_stat = [&](int i) -> float {
      const int q_first = (i_start + i) * 128;
      float v = tid < 64 ? CUDART_INF_F : 0.f;  // rows past N: P = exp2(-inf) = 0, D = 0
      if (i < n && q_first + wg * 64 + (tid & 63) < p.Nq) v = __ldg(stat_src + q_first);
      return v;
    };
    float stat_next = fetch_stat(0);
    const uint64_t scale_log2_2 = pack_f32x2(p.scale_log2, p.scale_log2), scale_2 = pack_f32x2(p.scale, p.scale);
#ifdef FA_BWD_TRACE
    const bool prof = p.prof != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
    long long tw_x = 0, t_p1 = 0, tw_y = 0, t_p2 = 0, t_top = 0, t_top2 = 0, t_begin = prof ? clock64() : 0;
#define FA_PCLK(var) long long var = prof ? clock64() : 0
#define FA_PDO(stmt) do { if (prof) { stmt; } } while (0)
#else
#define FA_PCLK(var) do {} while (0)
#define FA_PDO(stmt) do {} while (0)
#endif
    for (int i = 0; i < n; ++i) {
      const int q0 = (i_start + i) * 128 + wg * 64;  // first query column of this warpgroup's half
      float *ld = sLD + (wg * 2 + (i & 1)) * 128;
      FA_PCLK(ca);
      ld[tid] = stat_next * stat_coef;
      FA_PCLK(cb);
      stat_next = fetch_stat(i + 1);
      named_bar_sync(1 + wg, 128);
      FA_PDO(t_top += cb - ca; t_top2 += clock64() - cb);
      // ---- phase 1: P^T = exp2(S^T * c - L * log2e) ----
      FA_PCLK(c0);
      mbar_wait(x_full, i & 1);
      FA_PCLK(c1);
      tc_fence_after();
      uint32_t pr[2][32];  // S^T, then P^T (fp32 bits), kept for phase 2
      tmem_ld32(tX, pr[0]);
      tmem_ld32(tX + 32, pr[1]);
      tmem_wait_ld();
      const bool diag = p.causal && (q0 < key0 + 128);
      const uint32_t ld_s = smem_u32(ld);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
        if (!diag) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            uint64_t la, lb;  // -L*log2e of four consecutive query columns
            lds_v2b64(ld_s + (c * 32 + e) * 4, la, lb);
            const uint64_t xa = fma_f32x2(pack_u32x2(pr[c][e], pr[c][e + 1]), scale_log2_2, la);
            const uint64_t xb = fma_f32x2(pack_u32x2(pr[c][e + 2], pr[c][e + 3]), scale_log2_2, lb);
            // both warpgroups run this phase at the same time, two warps per scheduler sharing
            // one MUFU unit: a share of the pairs goes to the FMA pipe instead
            const uint64_t pa = exp2_pair(xa, emulate_pair(e >> 1)), pb = exp2_pair(xb, emulate_pair((e >> 1) + 1));
            pr[c][e] = __float_as_uint(lo_f32(pa)); pr[c][e + 1] = __float_as_uint(hi_f32(pa));
            pr[c][e + 2] = __float_as_uint(lo_f32(pb)); pr[c][e + 3] = __float_as_uint(hi_f32(pb));
            pk[e >> 1] = pack2<IS_BF16>(lo_f32(pa), hi_f32(pa));
            pk[(e >> 1) + 1] = pack2<IS_BF16>(lo_f32(pb), hi_f32(pb));
          }
        } else {  // diagonal tile: keys after the query are masked
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            uint64_t la, lb;
            lds_v2b64(ld_s + (c * 32 + (e & ~3)) * 4, la, lb);
            const uint64_t x2 = fma_f32x2(pack_u32x2(pr[c][e], pr[c][e + 1]), scale_log2_2, (e & 2) ? lb : la);
            float p0 = ex2(lo_f32(x2)), p1 = ex2(hi_f32(x2));
            if (key > q0 + c * 32 + e) p0 = 0.f;
            if (key > q0 + c * 32 + e + 1) p1 = 0.f;
            pr[c][e] = __float_as_uint(p0);
            pr[c][e + 1] = __float_as_uint(p1);
            pk[e >> 1] = pack2<IS_BF16>(p0, p1);
          }
        }
        tmem_st16(tX + c * 16, pk);
        if (kDkdvParts == 2 || c == 1) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive_warp(&p_ready[kDkdvParts == 2 ? c : 0]);
        }
      }
      // ---- phase 2: dS^T = P^T o (dP^T * scale - D * scale) ----
      FA_PCLK(c2);
      mbar_wait(y_full, i & 1);
      FA_PCLK(c3);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t y[32], dk[16];
        tmem_ld32(tY + c * 32, y);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          uint64_t da, db;  // -D*scale of four consecutive query columns
          lds_v2b64(ld_s + (64 + c * 32 + e) * 4, da, db);
          const uint64_t ga = fma_f32x2(pack_u32x2(y[e], y[e + 1]), scale_2, da);
          const uint64_t gb = fma_f32x2(pack_u32x2(y[e + 2], y[e + 3]), scale_2, db);
          const uint64_t d2a = mul_f32x2(pack_u32x2(pr[c][e], pr[c][e + 1]), ga);
          const uint64_t d2b = mul_f32x2(pack_u32x2(pr[c][e + 2], pr[c][e + 3]), gb);
          dk[e >> 1] = pack2<IS_BF16>(lo_f32(d2a), hi_f32(d2a));
          dk[(e >> 1) + 1] = pack2<IS_BF16>(lo_f32(d2b), hi_f32(d2b));
        }
        tmem_st16(tY + c * 16, dk);  // columns [16c, 16c+16) of my half were read in chunk <= c
        if (kDkdvParts == 2 || c == 1) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive_warp(&ds_ready[kDkdvParts == 2 ? c : 0]);
        }
      }
      FA_PDO(tw_x += c1 - c0; t_p1 += c2 - c1; tw_y += c3 - c2; t_p2 += clock64() - c3);
    }
#ifdef FA_BWD_TRACE
    if (prof) {
      long long *o = p.prof + wg * 8;
      o[0] = n; o[1] = tw_x; o[2] = t_p1; o[3] = tw_y; o[4] = t_p2; o[5] = clock64() - t_begin; o[6] = t_top; o[7] = t_top2;
    }
#endif
#undef FA_PCLK
#undef FA_PDO
    // ------------------------------ epilogue ------------------------------
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t tAcc = tmem_base + lane_off + 256 + wg * D;  // wg 0 -> dV, wg 1 -> dK
    float *dst = (wg == 0 ? p.dV : p.dK) + kv_off + (int64_t)key * D;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t a[32];
      tmem_ld32(tAcc + c * 32, a);
      tmem_wait_ld();
      if (key < p.Nk) {
        float4 *d4 = reinterpret_cast<float4 *>(dst + c * 32);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          d4[e] = make_float4(__uint_as_float(a[4 * e]), __uint_as_float(a[4 * e + 1]),
                              __uint_as_float(a[4 * e + 2]), __uint_as_float(a[4 * e + 3]));
      }
    }
  } else {
    setmaxnreg_dec<64>();
    if (warp == kLoadWarp) {
      // ============================ TMA producer ============================
      if (elect_one()) {
        prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmdO);
        mbar_arrive_expect_tx(res_full, 2 * Cfg::kTile);
        tma_load_tile<D>(sK, &tmK, res_full, key0, h, b);
        tma_load_tile<D>(sV, &tmV, res_full, key0, h, b);
        for (int i = 0; i < n; ++i) {
          const int q0 = (i_start + i) * 128;
          const int qs = i % Cfg::kQSlots, ds = i % Cfg::kDoSlots;
          mbar_wait(&q_empty[qs], ((i / Cfg::kQSlots) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[qs], Cfg::kTile);
          tma_load_tile<D>(sQ + qs * Cfg::kTile, &tmQ, &q_full[qs], q0, h, b);
          mbar_wait(&do_empty[ds], ((i / Cfg::kDoSlots) & 1) ^ 1);
          mbar_arrive_expect_tx(&do_full[ds], Cfg::kTile);
          tma_load_tile<D>(sDO + ds * Cfg::kTile, &tmdO, &do_full[ds], q0, h, b);
        }
      }
      __syncwarp();
    } else if (warp == kMmaWarp) {
      // ============================= MMA issuer =============================
      if (elect_one()) {
        constexpr uint32_t idesc_xy = make_idesc(128, 128, IS_BF16, 0, 0);
        constexpr uint32_t idesc_acc = make_idesc(128, D, IS_BF16, 0, 1);
        const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV), sQ_a = smem_u32(sQ), sDO_a = smem_u32(sDO);
        const uint32_t tX = tmem_base, tY = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 256 + D;
        auto q_addr = [&](int i) { return sQ_a + (i % Cfg::kQSlots) * Cfg::kTile; };
        auto do_addr = [&](int i) { return sDO_a + (i % Cfg::kDoSlots) * Cfg::kTile; };
        auto issue_x = [&](int i) {  // S^T = K Q_i^T
          mbar_wait(&q_full[i % Cfg::kQSlots], (i / Cfg::kQSlots) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)
            mma_ss(tX, kmajor_desc(sK_a, kk), kmajor_desc(q_addr(i), kk), idesc_xy, kk > 0);
          tc_commit(x_full);
        };
        auto issue_y = [&](int i) {  // dP^T = V dO_i^T
          mbar_wait(&do_full[i % Cfg::kDoSlots], (i / Cfg::kDoSlots) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)
            mma_ss(tY, kmajor_desc(sV_a, kk), kmajor_desc(do_addr(i), kk), idesc_xy, kk > 0);
          tc_commit(y_full);
        };
#ifdef FA_BWD_TRACE
        const bool mprof = p.prof != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0;
        long long mw_p = 0, mw_ds = 0, m_begin = mprof ? clock64() : 0, w0 = 0;
#define FA_MCLK() do { if (mprof) w0 = clock64(); } while (0)
#define FA_MACC(acc) do { if (mprof) acc += clock64() - w0; } while (0)
#else
#define FA_MCLK() do {} while (0)
#define FA_MACC(acc) do {} while (0)
#endif
        mbar_wait(res_full, 0);
        tc_fence_after();
        issue_x(0);
        issue_y(0);
        for (int i = 0; i < n; ++i) {
          FA_MCLK();
          // dV += P^T dO_i (K = 128 query rows); part c = columns [32c, 32c+32) of both warpgroups'
          // halves = k-steps {2c, 2c+1, 4+2c, 5+2c}
#pragma unroll
          for (int part = 0; part < kDkdvParts; ++part) {
            mbar_wait(&p_ready[part], i & 1);
            tc_fence_after();
#pragma unroll
            for (int q = 0; q < 8 / kDkdvParts; ++q) {
              const int kk = kDkdvParts == 2 ? (q >> 1) * 4 + part * 2 + (q & 1) : q;
              mma_ts(tdV, tX + (kk >> 2) * 64 + (kk & 3) * 8, mnmajor_desc(do_addr(i), kk), idesc_acc,
                     (i > 0 || part > 0 || q > 0) ? 1u : 0u);
            }
          }
          FA_MACC(mw_p);
          tc_commit(&do_empty[i % Cfg::kDoSlots]);
          if (i + 1 < n) issue_x(i + 1);
          FA_MCLK();
#pragma unroll
          for (int part = 0; part < kDkdvParts; ++part) {  // dK += dS^T Q_i
            mbar_wait(&ds_ready[part], i & 1);
            tc_fence_after();
#pragma unroll
            for (int q = 0; q < 8 / kDkdvParts; ++q) {
              const int kk = kDkdvParts == 2 ? (q >> 1) * 4 + part * 2 + (q & 1) : q;
              mma_ts(tdK, tY + (kk >> 2) * 64 + (kk & 3) * 8, mnmajor_desc(q_addr(i), kk), idesc_acc,
                     (i > 0 || part > 0 || q > 0) ? 1u : 0u);
            }
          }
          FA_MACC(mw_ds);
          tc_commit(&q_empty[i % Cfg::kQSlots]);
          if (i == n - 1) tc_commit(acc_full);
          if (i + 1 < n) issue_y(i + 1);
        }
#ifdef FA_BWD_TRACE
        if (mprof) { p.prof[16] = mw_p; p.prof[17] = mw_ds; p.prof[18] = clock64() - m_begin; }
#endif
#undef FA_MCLK
#undef FA_MACC
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------
// 3. dQ : CTA owns 2 x 128 query rows.  Warpgroup t owns tile t (one thread per query row, all
//    128 key columns); the two tiles take turns on ONE S / dP buffer pair:
//    TMEM: X = S [0,128)   Y = dP [128,256) (dS aliases its first 64 columns)
//          dQ_0 [256,256+D)   dQ_1 [256+D,256+2D)
//    Work items k = (s, t) in order; per item X(k) -> Y(k) -> dQ(k).  The warpgroup releases X as
//    soon as it has copied it to registers, so the next item's S is computed under this item's
//    exponentials:   X(0) Y(0) | xc(0)? X(1) | dS(0)? dQ(0) Y(1) | xc(1)? X(2) | dS(1)? dQ(1) Y(2) ...
// ---------------------------------------------------------------------------
template <int D>
struct DqCfg : BwdCfg<D> {
  static constexpr int kKSlots = D == 128 ? 2 : 4;
  static constexpr int kVSlots = D == 128 ? 1 : 2;
  static constexpr int kSmemTiles = (4 + kKSlots + kVSlots) * BwdCfg<D>::kTile;
  static constexpr int kSmemBytes = kSmemTiles + 1024 + 256;
};

template <int D, int IS_BF16>
__global__ void __launch_bounds__(kBwdThreads, 1)
bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
              const BwdParams p) {
  using Cfg = DqCfg<D>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char *sRes = smem;                                   // [tile][Q | dO]
  unsigned char *sKs = smem + 4 * Cfg::kTile;                   // [kKSlots]
  unsigned char *sVs = sKs + Cfg::kKSlots * Cfg::kTile;         // [kVSlots]
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemTiles);
  uint64_t *res_full = bars;        // [2]
  uint64_t *acc_full = bars + 2;    // [2]
  uint64_t *x_full = bars + 4;      // [2] S of tile t ready
  uint64_t *y_full = bars + 6;      // [2] dP of tile t ready
  uint64_t *x_taken = bars + 8;     // [2] warpgroup t has copied S to registers
  uint64_t *ds_ready = bars + 10;   // [2][4] warpgroup t has stored part c of dS
  uint64_t *k_full = bars + 18;
  uint64_t *k_empty = k_full + Cfg::kKSlots;
  uint64_t *v_full = k_empty + Cfg::kKSlots;
  uint64_t *v_empty = v_full + Cfg::kVSlots;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(v_empty + Cfg::kVSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const BlockCoord bc = decode_block(p.group, p.n_heads, p.H);
  if (bc.b < 0) return;
  const int h = bc.h, b = bc.b;
  const int qb = p.causal ? ((p.Nq + 255) / 256 - 1 - bc.blk) : bc.blk;  // heaviest first (sched.cuh)
  const int q_row0 = qb * 256;
  const int n_tiles_all = (p.Nk + 127) / 128;
  int n_t[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int r0 = q_row0 + t * 128;
    n_t[t] = r0 >= p.Nq ? 0 : (p.causal ? min(n_tiles_all, r0 / 128 + 1) : n_tiles_all);
  }
  const int nmax = max(n_t[0], n_t[1]);
  const int64_t head_off = (int64_t)b * p.batch_stride + (int64_t)h * p.head_stride;
  const int64_t vec_off = head_off / D;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&res_full[i], 1); mbar_init(&acc_full[i], 1);
      mbar_init(&x_full[i], 1); mbar_init(&y_full[i], 1);
      mbar_init(&x_taken[i], 4 * kArrivalsPerWarp);
      for (int c = 0; c < 4; ++c) mbar_init(&ds_ready[4 * i + c], 4 * kArrivalsPerWarp);
    }
    for (int i = 0; i < Cfg::kKSlots; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < Cfg::kVSlots; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ======================= element-wise warpgroups =======================
    setmaxnreg_inc<216>();
    const int t = warp >> 2;
    const int tid = (warp & 3) * 32 + lane;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tX = tmem_base + lane_off;
    const uint32_t tY = tmem_base + lane_off + 128;
    const int row = q_row0 + t * 128 + tid;
    const int nt = n_t[t];
    const float l2 = row < p.Nq ? __ldg(p.L + vec_off + row) * kLog2e : CUDART_INF_F;
    const float dl = row < p.Nq ? __ldg(p.delta + vec_off + row) : 0.f;
    const uint64_t scale_log2_2 = pack_f32x2(p.scale_log2, p.scale_log2), scale_2 = pack_f32x2(p.scale, p.scale);
    const uint64_t neg_l2_2 = pack_f32x2(-l2, -l2), neg_dls_2 = pack_f32x2(-dl * p.scale, -dl * p.scale);
    for (int s = 0; s < nt; ++s) {
      // ---- phase 1: copy S out (frees X for the other tile), P = exp2(S * c - L * log2e) ----
      mbar_wait(&x_full[t], s & 1);
      tc_fence_after();
      FA_BTRACE(tid == 0, s, t * 8 + 0);
      uint32_t pr[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tX + c * 32, pr[c]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive_warp(&x_taken[t]);
      FA_BTRACE(tid == 0, s, t * 8 + 1);
      const int k0 = s * 128;
      const bool diag = p.causal && (s == nt - 1);
      if (!diag) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const uint64_t x2 = fma_f32x2(pack_u32x2(pr[c][e], pr[c][e + 1]), scale_log2_2, neg_l2_2);
            const uint64_t p2 = exp2_pair(x2, emulate_pair<FA_BWD_EMU_DQ>(e >> 1));
            pr[c][e] = __float_as_uint(lo_f32(p2));
            pr[c][e + 1] = __float_as_uint(hi_f32(p2));
          }
      } else {  // diagonal tile: keys after the query are masked
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const uint64_t x2 = fma_f32x2(pack_u32x2(pr[c][e], pr[c][e + 1]), scale_log2_2, neg_l2_2);
            float p0 = ex2(lo_f32(x2)), p1 = ex2(hi_f32(x2));
            if (k0 + c * 32 + e > row) p0 = 0.f;
            if (k0 + c * 32 + e + 1 > row) p1 = 0.f;
            pr[c][e] = __float_as_uint(p0);
            pr[c][e + 1] = __float_as_uint(p1);
          }
      }
      // ---- phase 2: dS = P o (dP * scale - D * scale), 16-bit, over the first half of Y ----
      FA_BTRACE(tid == 0, s, t * 8 + 2);
      mbar_wait(&y_full[t], s & 1);
      tc_fence_after();
      FA_BTRACE(tid == 0, s, t * 8 + 3);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t y[32], dk[16];
        tmem_ld32(tY + c * 32, y);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const uint64_t g2 = fma_f32x2(pack_u32x2(y[e], y[e + 1]), scale_2, neg_dls_2);
          const uint64_t d2 = mul_f32x2(pack_u32x2(pr[c][e], pr[c][e + 1]), g2);
          dk[e >> 1] = pack2<IS_BF16>(lo_f32(d2), hi_f32(d2));
        }
        tmem_st16(tY + c * 16, dk);  // columns [16c, 16c+16) were read in chunk <= c
        if (kDqParts == 4 || c == 3) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive_warp(&ds_ready[4 * t + (kDqParts == 4 ? c : 0)]);
          FA_BTRACE(tid == 0 && (c == 0 || c == 3), s, t * 8 + 4 + (c == 3));
        }
      }
    }
    if (nt > 0) {
      mbar_wait(&acc_full[t], 0);
      tc_fence_after();
      const uint32_t tAcc = tmem_base + lane_off + 256 + t * D;
      float *dst = p.dQ + head_off + (int64_t)row * D;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t a[32];
        tmem_ld32(tAcc + c * 32, a);
        tmem_wait_ld();
        if (row < p.Nq) {
          float4 *d4 = reinterpret_cast<float4 *>(dst + c * 32);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float4 v = make_float4(__uint_as_float(a[4 * e]), __uint_as_float(a[4 * e + 1]),
                                   __uint_as_float(a[4 * e + 2]), __uint_as_float(a[4 * e + 3]));
            if (p.acc_dq) {  // single owner per element: a plain read-add-write is deterministic
              const float4 o = d4[e];
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            d4[e] = v;
          }
        }
      }
    }
  } else {
    setmaxnreg_dec<64>();
    if (warp == kLoadWarp) {
      if (elect_one()) {
        prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmdO);
#pragma unroll
        for (int t = 0; t < 2; ++t)
          if (n_t[t] > 0) {
            mbar_arrive_expect_tx(&res_full[t], 2 * Cfg::kTile);
            tma_load_tile<D>(sRes + t * 2 * Cfg::kTile, &tmQ, &res_full[t], q_row0 + t * 128, h, b);
            tma_load_tile<D>(sRes + t * 2 * Cfg::kTile + Cfg::kTile, &tmdO, &res_full[t], q_row0 + t * 128, h, b);
          }
        for (int s = 0; s < nmax; ++s) {
          const int ks = s % Cfg::kKSlots, vs = s % Cfg::kVSlots;
          mbar_wait(&k_empty[ks], ((s / Cfg::kKSlots) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[ks], Cfg::kTile);
          tma_load_tile<D>(sKs + ks * Cfg::kTile, &tmK, &k_full[ks], s * 128, h, b);
          mbar_wait(&v_empty[vs], ((s / Cfg::kVSlots) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[vs], Cfg::kTile);
          tma_load_tile<D>(sVs + vs * Cfg::kTile, &tmV, &v_full[vs], s * 128, h, b);
        }
      }
      __syncwarp();
    } else if (warp == kMmaWarp) {
      if (elect_one()) {
        constexpr uint32_t idesc_xy = make_idesc(128, 128, IS_BF16, 0, 0);
        constexpr uint32_t idesc_acc = make_idesc(128, D, IS_BF16, 0, 1);
        const uint32_t sRes_a = smem_u32(sRes), sK_a = smem_u32(sKs), sV_a = smem_u32(sVs);
        const uint32_t tX = tmem_base, tY = tmem_base + 128;
        auto k_addr = [&](int s) { return sK_a + (s % Cfg::kKSlots) * Cfg::kTile; };
        auto v_addr = [&](int s) { return sV_a + (s % Cfg::kVSlots) * Cfg::kTile; };
        // item k -> (s, t); items run s-major over the tiles that still have work
        auto item_s = [&](int k) { return k < 2 * min(n_t[0], n_t[1]) ? k >> 1 : k - min(n_t[0], n_t[1]); };
        auto item_t = [&](int k) { return k < 2 * min(n_t[0], n_t[1]) ? k & 1 : (n_t[0] > n_t[1] ? 0 : 1); };
        const int n_items = n_t[0] + n_t[1];
        int k_waited = -1, v_waited = -1;  // highest s whose K / V tile is known to have landed
        bool res_waited[2] = {false, false};
        auto issue_x = [&](int k) {  // S = Q_t K_s^T
          const int s = item_s(k), t = item_t(k);
          if (!res_waited[t]) { mbar_wait(&res_full[t], 0); res_waited[t] = true; }
          if (s > k_waited) { mbar_wait(&k_full[s % Cfg::kKSlots], (s / Cfg::kKSlots) & 1); k_waited = s; }
          tc_fence_after();
          const uint32_t q_a = sRes_a + t * 2 * Cfg::kTile;
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)
            mma_ss(tX, kmajor_desc(q_a, kk), kmajor_desc(k_addr(s), kk), idesc_xy, kk > 0);
          tc_commit(&x_full[t]);
        };
        auto issue_y = [&](int k) {  // dP = dO_t V_s^T
          const int s = item_s(k), t = item_t(k);
          if (s > v_waited) { mbar_wait(&v_full[s % Cfg::kVSlots], (s / Cfg::kVSlots) & 1); v_waited = s; }
          tc_fence_after();
          const uint32_t do_a = sRes_a + t * 2 * Cfg::kTile + Cfg::kTile;
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)
            mma_ss(tY, kmajor_desc(do_a, kk), kmajor_desc(v_addr(s), kk), idesc_xy, kk > 0);
          FA_BTRACE(true, s, 32 + t);
          tc_commit(&y_full[t]);
          FA_BTRACE(true, s, 34 + t);
          // last use of V_s: release its slot once this MMA has completed
          if (k + 1 >= n_items || item_s(k + 1) != s) tc_commit(&v_empty[s % Cfg::kVSlots]);
        };
        if (n_items > 0) {
          issue_x(0);
          issue_y(0);
        }
        for (int k = 0; k < n_items; ++k) {
          const int s = item_s(k), t = item_t(k);
          FA_BTRACE(true, s, 16 + t * 8 + 0);
          if (k + 1 < n_items) {
            mbar_wait(&x_taken[t], s & 1);  // X is free again
            FA_BTRACE(true, s, 16 + t * 8 + 1);
            issue_x(k + 1);
            FA_BTRACE(true, s, 16 + t * 8 + 2);
          }
#pragma unroll
          for (int part = 0; part < kDqParts; ++part) {  // dQ_t += dS K_s   (K = 128 keys)
            mbar_wait(&ds_ready[4 * t + part], s & 1);
            tc_fence_after();
            FA_BTRACE(part == 0 || part == kDqParts - 1, s, 16 + t * 8 + 3 + (part > 0));
#pragma unroll
            for (int kk = part * (8 / kDqParts); kk < (part + 1) * (8 / kDqParts); ++kk)
              mma_ts(tmem_base + 256 + t * D, tY + kk * 8, mnmajor_desc(k_addr(s), kk), idesc_acc,
                     (s > 0 || kk > 0) ? 1u : 0u);
          }
          if (s == n_t[t] - 1) tc_commit(&acc_full[t]);
          if (k + 1 >= n_items || item_s(k + 1) != s) tc_commit(&k_empty[s % Cfg::kKSlots]);
          FA_BTRACE(true, s, 16 + t * 8 + 5);
          if (k + 1 < n_items) issue_y(k + 1);
          FA_BTRACE(true, s, 16 + t * 8 + 6);
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int D, int IS_BF16>
int configure_bwd() {
  static DeviceOnce configured;  // the attribute is per device
  return configured.run([] {
    FA_CUDA_CHECK(cudaFuncSetAttribute(bwd_dkdv_kernel<D, IS_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       DkdvCfg<D>::kSmemBytes));
    FA_CUDA_CHECK(cudaFuncSetAttribute(bwd_dq_kernel<D, IS_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       DqCfg<D>::kSmemBytes));
    cudaFuncAttributes attr;  // forces the (lazily loaded) kernels into the context
    FA_CUDA_CHECK(cudaFuncGetAttributes(&attr, bwd_dkdv_kernel<D, IS_BF16>));
    FA_CUDA_CHECK(cudaFuncGetAttributes(&attr, bwd_dq_kernel<D, IS_BF16>));
    FA_CUDA_CHECK(cudaFuncGetAttributes(&attr, bwd_delta_kernel<D, IS_BF16>));
    return (int)FA_OK;
  });
}

template <int D, int IS_BF16>
int launch_bwd_impl(const CUtensorMap *const *maps, const BwdParams &p, int B, cudaStream_t stream) {
  const int rc = configure_bwd<D, IS_BF16>();
  if (rc != FA_OK) return rc;
  // maps: [0] Q, [1] K, [2] V, [3] dO, all with 128-row boxes
  BwdParams q = p;
  q.n_heads = B * p.H;
  q.group = dispatch_group(p.causal != 0, (int64_t)2 * (p.Nq > p.Nk ? p.Nq : p.Nk) * D * 2, q.n_heads);
  if ((p.Nk + 127) / 128 > 65535 || (p.Nq + 255) / 256 > 65535) q.group = 1;  // grid.y limit of the grouped form
  if (p.dK != nullptr) {
    bwd_dkdv_kernel<D, IS_BF16><<<dispatch_grid(q.group, (p.Nk + 127) / 128, p.H, B), kBwdThreads, DkdvCfg<D>::kSmemBytes, stream>>>(
        *maps[0], *maps[1], *maps[2], *maps[3], q);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  if (p.dQ != nullptr) {
    bwd_dq_kernel<D, IS_BF16><<<dispatch_grid(q.group, (p.Nq + 255) / 256, p.H, B), kBwdThreads, DqCfg<D>::kSmemBytes, stream>>>(
        *maps[0], *maps[1], *maps[2], *maps[3], q);
    FA_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  return FA_OK;
}

}  // namespace

int preload_bwd_tc() {
  int rc;
  if ((rc = configure_bwd<64, 0>()) || (rc = configure_bwd<64, 1>()) || (rc = configure_bwd<128, 0>()) || (rc = configure_bwd<128, 1>()))
    return rc;
  return FA_OK;
}

// D_i = rowsum(O o dO) into `delta` ([B, H, Nq] addressed like L: offset / D + row)
int launch_bwd_delta(const void *O, const void *dO, float *delta, int Nq, int D, int64_t batch_stride,
                     int64_t head_stride, int B, int H, int dtype, cudaStream_t stream) {
  const dim3 grid((Nq + 7) / 8, H, B);
  const uint16_t *o = reinterpret_cast<const uint16_t *>(O), *g = reinterpret_cast<const uint16_t *>(dO);
  if (D == 64) {
    if (dtype == FA_DTYPE_BF16) bwd_delta_kernel<64, 1><<<grid, 256, 0, stream>>>(o, g, delta, Nq, H, batch_stride, head_stride);
    else bwd_delta_kernel<64, 0><<<grid, 256, 0, stream>>>(o, g, delta, Nq, H, batch_stride, head_stride);
  } else {
    if (dtype == FA_DTYPE_BF16) bwd_delta_kernel<128, 1><<<grid, 256, 0, stream>>>(o, g, delta, Nq, H, batch_stride, head_stride);
    else bwd_delta_kernel<128, 0><<<grid, 256, 0, stream>>>(o, g, delta, Nq, H, batch_stride, head_stride);
  }
  FA_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return FA_OK;
}

// Rectangular backward block: gradients of attention(Q[Nq], K[Nk], V[Nk]) given L and delta of the
// FULL softmax rows.  dK/dV (both or neither) are overwritten; dQ is overwritten or accumulated.
int launch_bwd_tc_rect(const void *Q, const void *K, const void *V, const void *dO, const float *L,
                       const float *delta, float *dQ, float *dK, float *dV, int Nq, int Nk, int D, float scale,
                       int64_t q_batch_stride, int64_t q_head_stride, int64_t kv_batch_stride,
                       int64_t kv_head_stride, int is_causal, int acc_dq, int B, int H, int dtype,
                       cudaStream_t stream, void *fused_sems) {
  FA_REQUIRE(Q && K && V && dO && L && delta, "null tensor pointer");
  FA_REQUIRE((dK == nullptr) == (dV == nullptr), "dK and dV go together");
  FA_REQUIRE(Nq >= 1 && Nk >= 1, "N must be >= 1 (got %d x %d)", Nq, Nk);
  FA_REQUIRE(!is_causal || Nq == Nk, "causal attention needs Nq == Nk");
  FA_REQUIRE(D == 64 || D == 128, "D must be 64 or 128 (got %d)", D);
  FA_REQUIRE(B >= 1 && H >= 1 && H <= 65535 && B <= 65535, "bad B/H (%d, %d)", B, H);
  FA_REQUIRE(dtype == FA_DTYPE_FP16 || dtype == FA_DTYPE_BF16, "dtype must be FA_DTYPE_FP16 or FA_DTYPE_BF16");
  FA_REQUIRE(scale > 0.f, "scale must be positive");
  FA_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(dO) && aligned16(dQ) && aligned16(dK) &&
                 aligned16(dV),
             "tensors must be 16-byte aligned");
  FA_REQUIRE(q_batch_stride % D == 0 && q_head_stride % D == 0 && kv_batch_stride % 8 == 0 && kv_head_stride % 8 == 0,
             "Q-side strides must be multiples of D, K/V-side strides multiples of 8");
  FA_REQUIRE((H == 1 || (q_head_stride >= (int64_t)Nq * D && kv_head_stride >= (int64_t)Nk * D)) &&
                 (B == 1 || (q_batch_stride >= (int64_t)Nq * D && kv_batch_stride >= (int64_t)Nk * D)),
             "heads overlap: stride smaller than N*D");
  // Default: the two-kernel form below (seven GEMMs, every gradient tile owned by one CTA) -- the faster
  // one on B200 for every shape that fills the GPU.  fa_set_backward_algorithm(
  // FA_BWD_FUSED) selects the single fused kernel (five GEMMs per tile pair, dQ by ordered TMA
  // add-reduction, bwd_fused.cu) when all three gradients are wanted and the workspace holds its counters.
  if (bwd_mode() == FA_BWD_FUSED && fused_sems != nullptr && dQ != nullptr && dK != nullptr)
    return launch_bwd_fused(Q, K, V, dO, L, delta, dQ, dK, dV, Nq, Nk, D, scale, q_batch_stride, q_head_stride,
                            kv_batch_stride, kv_head_stride, is_causal, acc_dq, B, H, dtype, fused_sems, stream);
  const CUtensorMap *maps[4];
  int rc;
  if ((rc = tensor_map_bhnd(&maps[0], Q, dtype, Nq, D, H, B, q_head_stride, q_batch_stride, 128)) != FA_OK) return rc;
  if ((rc = tensor_map_bhnd(&maps[1], K, dtype, Nk, D, H, B, kv_head_stride, kv_batch_stride, 128)) != FA_OK) return rc;
  if ((rc = tensor_map_bhnd(&maps[2], V, dtype, Nk, D, H, B, kv_head_stride, kv_batch_stride, 128)) != FA_OK) return rc;
  if ((rc = tensor_map_bhnd(&maps[3], dO, dtype, Nq, D, H, B, q_head_stride, q_batch_stride, 128)) != FA_OK) return rc;
  BwdParams p = {};
  p.L = L;
  p.delta = delta;
  p.dQ = dQ; p.dK = dK; p.dV = dV;
  p.Nq = Nq; p.Nk = Nk; p.H = H;
  p.scale = scale;
  p.scale_log2 = scale * kLog2e;
  p.batch_stride = q_batch_stride;
  p.head_stride = q_head_stride;
  p.kv_batch_stride = kv_batch_stride;
  p.kv_head_stride = kv_head_stride;
  p.causal = is_causal ? 1 : 0;
  p.acc_dq = acc_dq ? 1 : 0;
#ifdef FA_BWD_TRACE
  p.prof = g_trace_buffer;
#endif
  if (D == 64)
    return dtype == FA_DTYPE_BF16 ? launch_bwd_impl<64, 1>(maps, p, B, stream) : launch_bwd_impl<64, 0>(maps, p, B, stream);
  return dtype == FA_DTYPE_BF16 ? launch_bwd_impl<128, 1>(maps, p, B, stream) : launch_bwd_impl<128, 0>(maps, p, B, stream);
}

int launch_bwd_tc(const void *Q, const void *K, const void *V, const void *O, const void *dO,
                  const float *L, float *dQ, float *dK, float *dV, int N, int D, float scale,
                  int64_t batch_stride, int64_t head_stride, int is_causal, int B, int H, int dtype,
                  void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  FA_REQUIRE(Q && K && V && O && dO && L && dQ && dK && dV, "null tensor pointer");
  FA_REQUIRE(N >= 1, "N must be >= 1 (got %d)", N);
  FA_REQUIRE(D == 64 || D == 128, "D must be 64 or 128 (got %d)", D);
  FA_REQUIRE(B >= 1 && H >= 1 && H <= 65535 && B <= 65535, "bad B/H (%d, %d)", B, H);
  FA_REQUIRE(aligned16(O), "tensors must be 16-byte aligned");
  FA_REQUIRE(batch_stride % D == 0 && head_stride % D == 0, "strides must be multiples of D");
  // workspace = [ordering counters of the fused kernel | delta]; delta is indexed like L: offset / D + row,
  // so its extent follows the strides
  const size_t need = fa_workspace_bytes_backward(N, D, B, H);
  const size_t sem_bytes = bwd_fused_sem_bytes(N, B, H);
  const int64_t last = ((int64_t)(B - 1) * batch_stride + (int64_t)(H - 1) * head_stride) / D + N;
  if (workspace == nullptr || workspace_bytes < need || sem_bytes + (size_t)last * sizeof(float) > workspace_bytes)
    return set_error(FA_ERR_WORKSPACE,
                     "backward workspace too small: need %zu bytes (fa_workspace_bytes_backward; contiguous "
                     "[B,H,N,D] layout assumed), got %zu",
                     need, workspace_bytes);
  FA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "workspace must be 16-byte aligned");
  float *delta = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + sem_bytes);
  int rc = launch_bwd_delta(O, dO, delta, N, D, batch_stride, head_stride, B, H, dtype, stream);
  if (rc != FA_OK) return rc;
  return launch_bwd_tc_rect(Q, K, V, dO, L, delta, dQ, dK, dV, N, N, D, scale, batch_stride, head_stride, batch_stride,
                            head_stride, is_causal, 0, B, H, dtype, stream, workspace);
}

}  // namespace fa
