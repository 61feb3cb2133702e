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
//   2. bwd_dkdv_kernel  : one CTA per 128-key tile, streams 64-row Q/dO half tiles,
//                         S^T = K Q^T and dP^T = V dO^T land transposed in TMEM so that
//                         P^T and dS^T are directly the TMEM A operands of
//                         dV += P^T dO and dK += dS^T Q  (4 GEMMs per tile pair)
//   3. bwd_dq_kernel    : one CTA per 2 x 128 query rows, streams 64-row K/V half tiles,
//                         S = Q K^T, dP = dO V^T, dQ += dS K  (3 GEMMs per tile pair)
// S and dP are recomputed in both kernels (7 GEMMs instead of 5): that is the price of
// an atomic-free, order-independent dQ.
//
// Warp roles in both kernels: warps 0-3 and 4-7 are two element-wise warpgroups (one
// thread per TMEM lane), warp 8 issues MMAs, warp 9 drives TMA.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fa_internal.h"
#include "sm100_ptx.cuh"
#include "tensormap.h"

namespace fa {
namespace {

using namespace ptx;

constexpr int kBwdThreads = 320;
constexpr int kMmaWarp = 8;
constexpr int kLoadWarp = 9;
constexpr float kLog2e = 1.4426950408889634f;

struct BwdParams {
  const float *L;      // [B, H, N] log-sum-exp of the scaled scores (natural log)
  const float *delta;  // [B, H, N] D_i (workspace)
  float *dQ, *dK, *dV;
  int N, H;
  float scale, scale_log2;
  int64_t batch_stride, head_stride;  // elements
  int causal;
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

// Shared-memory geometry.  Every tile is a stack of [rows][64 elements] chunks (128-byte
// rows, 128-byte swizzle) as TMA writes them.
template <int D>
struct BwdCfg {
  static constexpr int kChunks = D / 64;
  static constexpr int kChunk128 = 128 * 128;  // bytes of a 128-row chunk
  static constexpr int kChunk64 = 64 * 128;    // bytes of a 64-row chunk
  static constexpr int kTile128 = kChunks * kChunk128;
  static constexpr int kTile64 = kChunks * kChunk64;
  static constexpr int kStageBytes = 2 * kTile64;  // a streamed pair of half tiles
};

// K-major operand, k-step kk (16 elements of the head dim): chunk kk/4, 32 bytes per step
template <int CHUNK_BYTES>
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t tile_addr, int kk) {
  return make_sdesc_sw128(tile_addr + (kk >> 2) * CHUNK_BYTES + (kk & 3) * 32, 16, 1024);
}
// MN-major B operand over a [rows][D] tile: k-step kk covers rows 16kk..16kk+15
template <int CHUNK_BYTES>
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t tile_addr, int kk) {
  return make_sdesc_sw128(tile_addr + kk * 2048, CHUNK_BYTES, 1024);
}

// ---------------------------------------------------------------------------
// 2. dK / dV : CTA owns keys [128 j, 128 j + 128); streams Q/dO half tiles of 64 rows.
//    TMEM: X_b = S^T  [b*64, +64)      (P^T  aliases its first 32 columns)
//          Y_b = dP^T [128 + b*64, +64) (dS^T aliases its first 32 columns)
//          dV [256, 256+D)   dK [256+D, 256+2D)
// ---------------------------------------------------------------------------
template <int D>
struct DkdvCfg : BwdCfg<D> {
  static constexpr int kStages = 4;
  static constexpr int kSmemTiles = 2 * BwdCfg<D>::kTile128 + kStages * BwdCfg<D>::kStageBytes;
  static constexpr int kSmemBytes = kSmemTiles + 2 * 2 * 128 * 4 /*L,D vectors*/ + 1024 + 256;
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
  unsigned char *sK = smem;
  unsigned char *sV = smem + Cfg::kTile128;
  unsigned char *sStage = smem + 2 * Cfg::kTile128;  // [stage][Q half | dO half]
  float *sLD = reinterpret_cast<float *>(smem + Cfg::kSmemTiles);  // [wg][slot][L2e 64 | delta 64]
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemTiles + 2 * 2 * 128 * 4);
  uint64_t *res_full = bars;          // [1]
  uint64_t *acc_full = bars + 1;      // [1]
  uint64_t *xy_full = bars + 2;       // [2]
  uint64_t *pds_full = bars + 4;      // [2]
  uint64_t *st_full = bars + 6;       // [kStages]
  uint64_t *st_empty = bars + 6 + Cfg::kStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6 + 2 * Cfg::kStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int key0 = j * 128;
  const int n_half_all = (p.N + 63) / 64;
  const int i_start = p.causal ? 2 * j : 0;  // first 64-row query half tile that sees these keys
  const int n = n_half_all - i_start;        // >= 1 because key0 < N
  const int64_t head_off = (int64_t)b * p.batch_stride + (int64_t)h * p.head_stride;
  const int64_t vec_off = head_off / D;

  if (threadIdx.x == 0) {
    mbar_init(res_full, 1);
    mbar_init(acc_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&xy_full[i], 1); mbar_init(&pds_full[i], 128); }
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ======================= element-wise warpgroups =======================
    const int wg = warp >> 2;
    const int tid = (warp & 3) * 32 + lane;  // TMEM lane = key row within the tile
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tX = tmem_base + lane_off + wg * 64;
    const uint32_t tY = tmem_base + lane_off + 128 + wg * 64;
    const int key = key0 + tid;
    // per-column statistics (L_i * log2e for threads 0-63, D_i * scale for threads 64-127) of a half
    // tile, fetched one iteration ahead so the global-memory latency hides behind the previous tile
    auto fetch_stat = [&](int s) -> float {
      const int qi = (i_start + s) * 64 + (tid & 63);
      if (s >= n || qi >= p.N) return tid < 64 ? CUDART_INF_F : 0.f;
      return tid < 64 ? __ldg(p.L + vec_off + qi) * kLog2e : __ldg(p.delta + vec_off + qi) * p.scale;
    };
    float stat_next = fetch_stat(wg);
    const uint64_t scale_log2_2 = pack_f32x2(p.scale_log2, p.scale_log2), scale_2 = pack_f32x2(p.scale, p.scale);
    int it = 0;
    for (int s = wg; s < n; s += 2, ++it) {
      const int q0 = (i_start + s) * 64;
      float *ld = sLD + (wg * 2 + (it & 1)) * 128;  // double-buffered per warpgroup
      ld[tid] = stat_next;
      stat_next = fetch_stat(s + 2);
      named_bar_sync(1 + wg, 128);
      mbar_wait(&xy_full[wg], it & 1);
      tc_fence_after();
      const bool diag = p.causal && (q0 < key0 + 128);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t x[32], y[32];
        tmem_ld32(tX + c * 32, x);
        tmem_ld32(tY + c * 32, y);
        tmem_wait_ld();
        uint32_t pp[16], ds[16];
        // packed fp32x2 math: P = exp2(x*c - L*log2e), dS = P * (y*scale - D*scale)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t l2 = *reinterpret_cast<const uint64_t *>(&ld[c * 32 + i]);
          const uint64_t dl = *reinterpret_cast<const uint64_t *>(&ld[64 + c * 32 + i]);
          const uint64_t e2 = fma_f32x2(pack_u32x2(x[i], x[i + 1]), scale_log2_2, neg_f32x2(l2));
          float p0 = ex2(lo_f32(e2)), p1 = ex2(hi_f32(e2));
          if (diag) {
            if (key > q0 + c * 32 + i) p0 = 0.f;
            if (key > q0 + c * 32 + i + 1) p1 = 0.f;
          }
          const uint64_t g2 = fma_f32x2(pack_u32x2(y[i], y[i + 1]), scale_2, neg_f32x2(dl));
          const uint64_t d2 = mul_f32x2(pack_f32x2(p0, p1), g2);
          pp[i >> 1] = pack2<IS_BF16>(p0, p1);
          ds[i >> 1] = pack2<IS_BF16>(lo_f32(d2), hi_f32(d2));
        }
        tmem_st16(tX + c * 16, pp);
        tmem_st16(tY + c * 16, ds);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&pds_full[wg]);
    }
    // ------------------------------ epilogue ------------------------------
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t tAcc = tmem_base + lane_off + 256 + wg * D;  // wg 0 -> dV, wg 1 -> dK
    float *dst = (wg == 0 ? p.dV : p.dK) + head_off + (int64_t)key * D;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t a[32];
      tmem_ld32(tAcc + c * 32, a);
      tmem_wait_ld();
      if (key < p.N) {
        float4 *d4 = reinterpret_cast<float4 *>(dst + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          d4[i] = make_float4(__uint_as_float(a[4 * i]), __uint_as_float(a[4 * i + 1]),
                              __uint_as_float(a[4 * i + 2]), __uint_as_float(a[4 * i + 3]));
      }
    }
  } else if (warp == kLoadWarp) {
    // ============================ TMA producer ============================
    if (elect_one()) {
      prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmdO);
      mbar_arrive_expect_tx(res_full, 2 * Cfg::kTile128);
#pragma unroll
      for (int c = 0; c < Cfg::kChunks; ++c) {
        tma_load_4d(sK + c * Cfg::kChunk128, &tmK, res_full, c * 64, key0, h, b);
        tma_load_4d(sV + c * Cfg::kChunk128, &tmV, res_full, c * 64, key0, h, b);
      }
      for (int s = 0; s < n; ++s) {
        const int stage = s % Cfg::kStages;
        mbar_wait(&st_empty[stage], ((s / Cfg::kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&st_full[stage], Cfg::kStageBytes);
        unsigned char *dst = sStage + stage * Cfg::kStageBytes;
        const int q0 = (i_start + s) * 64;
#pragma unroll
        for (int c = 0; c < Cfg::kChunks; ++c) {
          tma_load_4d(dst + c * Cfg::kChunk64, &tmQ, &st_full[stage], c * 64, q0, h, b);
          tma_load_4d(dst + Cfg::kTile64 + c * Cfg::kChunk64, &tmdO, &st_full[stage], c * 64, q0, h, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ============================= MMA issuer =============================
    if (elect_one()) {
      constexpr uint32_t idesc_xy = make_idesc(128, 64, IS_BF16, 0, 0);
      constexpr uint32_t idesc_acc = make_idesc(128, D, IS_BF16, 0, 1);
      const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV), sSt_a = smem_u32(sStage);
      auto issue_xy = [&](int s) {
        const int stage = s % Cfg::kStages, slot = s & 1;
        mbar_wait(&st_full[stage], (s / Cfg::kStages) & 1);
        tc_fence_after();
        const uint32_t q_a = sSt_a + stage * Cfg::kStageBytes, do_a = q_a + Cfg::kTile64;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)  // S^T = K Q^T
          mma_ss(tmem_base + slot * 64, kmajor_desc<Cfg::kChunk128>(sK_a, kk),
                 kmajor_desc<Cfg::kChunk64>(q_a, kk), idesc_xy, kk > 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)  // dP^T = V dO^T
          mma_ss(tmem_base + 128 + slot * 64, kmajor_desc<Cfg::kChunk128>(sV_a, kk),
                 kmajor_desc<Cfg::kChunk64>(do_a, kk), idesc_xy, kk > 0);
        tc_commit(&xy_full[slot]);
      };
      auto issue_acc = [&](int s) {
        const int stage = s % Cfg::kStages, slot = s & 1;
        mbar_wait(&pds_full[slot], (s >> 1) & 1);
        tc_fence_after();
        const uint32_t q_a = sSt_a + stage * Cfg::kStageBytes, do_a = q_a + Cfg::kTile64;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dV += P^T dO   (K = 64 query rows)
          mma_ts(tmem_base + 256, tmem_base + slot * 64 + kk * 8, mnmajor_desc<Cfg::kChunk64>(do_a, kk),
                 idesc_acc, (s > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dK += dS^T Q
          mma_ts(tmem_base + 256 + D, tmem_base + 128 + slot * 64 + kk * 8,
                 mnmajor_desc<Cfg::kChunk64>(q_a, kk), idesc_acc, (s > 0 || kk > 0) ? 1u : 0u);
        tc_commit(&st_empty[stage]);
        if (s == n - 1) tc_commit(acc_full);
      };
      mbar_wait(res_full, 0);
      tc_fence_after();
      issue_xy(0);
      if (n > 1) issue_xy(1);
      for (int s = 0; s < n; ++s) {
        issue_acc(s);
        if (s + 2 < n) issue_xy(s + 2);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------
// 3. dQ : CTA owns 2 x 128 query rows (two slots that ping-pong on the tensor core and
//    share every streamed K/V half tile of 64 keys).
//    TMEM: X_t = S [t*128, +64)   Y_t = dP [t*128+64, +64) (dS aliases its first 32 columns)
//          dQ_t [256 + t*D, +D)
// ---------------------------------------------------------------------------
template <int D>
struct DqCfg : BwdCfg<D> {
  static constexpr int kStages = D == 128 ? 3 : 4;
  static constexpr int kSmemTiles = 4 * BwdCfg<D>::kTile128 + kStages * BwdCfg<D>::kStageBytes;
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
  unsigned char *sRes = smem;                          // [slot][Q tile | dO tile]
  unsigned char *sStage = smem + 4 * Cfg::kTile128;    // [stage][K half | V half]
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemTiles);
  uint64_t *res_full = bars;       // [2]
  uint64_t *acc_full = bars + 2;   // [2]
  uint64_t *xy_full = bars + 4;    // [2]
  uint64_t *ds_full = bars + 6;    // [2]
  uint64_t *st_full = bars + 8;
  uint64_t *st_empty = bars + 8 + Cfg::kStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8 + 2 * Cfg::kStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int qb = p.causal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;  // heaviest first
  const int q_row0 = qb * 256;
  const int n_half_all = (p.N + 63) / 64;
  int n_t[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int r0 = q_row0 + t * 128;
    n_t[t] = r0 >= p.N ? 0 : (p.causal ? min(n_half_all, r0 / 64 + 2) : n_half_all);
  }
  const int nmax = max(n_t[0], n_t[1]);
  const int64_t head_off = (int64_t)b * p.batch_stride + (int64_t)h * p.head_stride;
  const int64_t vec_off = head_off / D;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&res_full[i], 1); mbar_init(&acc_full[i], 1);
      mbar_init(&xy_full[i], 1); mbar_init(&ds_full[i], 128);
    }
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ======================= element-wise warpgroups =======================
    const int t = warp >> 2;
    const int tid = (warp & 3) * 32 + lane;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tX = tmem_base + lane_off + t * 128;
    const uint32_t tY = tX + 64;
    const int row = q_row0 + t * 128 + tid;
    const int nt = n_t[t];
    const float l2 = row < p.N ? __ldg(p.L + vec_off + row) * kLog2e : CUDART_INF_F;
    const float dl = row < p.N ? __ldg(p.delta + vec_off + row) : 0.f;
    const uint64_t scale_log2_2 = pack_f32x2(p.scale_log2, p.scale_log2), scale_2 = pack_f32x2(p.scale, p.scale);
    const uint64_t neg_l2_2 = pack_f32x2(-l2, -l2), neg_dls_2 = pack_f32x2(-dl * p.scale, -dl * p.scale);
    for (int s = 0; s < nt; ++s) {
      mbar_wait(&xy_full[t], s & 1);
      tc_fence_after();
      const int k0 = s * 64;
      const bool diag = p.causal && (k0 + 63 > q_row0 + t * 128);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t x[32], y[32];
        tmem_ld32(tX + c * 32, x);
        tmem_ld32(tY + c * 32, y);
        tmem_wait_ld();
        uint32_t ds[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t e2 = fma_f32x2(pack_u32x2(x[i], x[i + 1]), scale_log2_2, neg_l2_2);
          float p0 = ex2(lo_f32(e2)), p1 = ex2(hi_f32(e2));
          if (diag) {
            if (k0 + c * 32 + i > row) p0 = 0.f;
            if (k0 + c * 32 + i + 1 > row) p1 = 0.f;
          }
          const uint64_t g2 = fma_f32x2(pack_u32x2(y[i], y[i + 1]), scale_2, neg_dls_2);
          const uint64_t d2 = mul_f32x2(pack_f32x2(p0, p1), g2);
          ds[i >> 1] = pack2<IS_BF16>(lo_f32(d2), hi_f32(d2));
        }
        tmem_st16(tY + c * 16, ds);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&ds_full[t]);
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
        if (row < p.N) {
          float4 *d4 = reinterpret_cast<float4 *>(dst + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            d4[i] = make_float4(__uint_as_float(a[4 * i]), __uint_as_float(a[4 * i + 1]),
                                __uint_as_float(a[4 * i + 2]), __uint_as_float(a[4 * i + 3]));
        }
      }
    }
  } else if (warp == kLoadWarp) {
    if (elect_one()) {
      prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmdO);
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (n_t[t] > 0) {
          mbar_arrive_expect_tx(&res_full[t], 2 * Cfg::kTile128);
          unsigned char *dst = sRes + t * 2 * Cfg::kTile128;
#pragma unroll
          for (int c = 0; c < Cfg::kChunks; ++c) {
            tma_load_4d(dst + c * Cfg::kChunk128, &tmQ, &res_full[t], c * 64, q_row0 + t * 128, h, b);
            tma_load_4d(dst + Cfg::kTile128 + c * Cfg::kChunk128, &tmdO, &res_full[t], c * 64, q_row0 + t * 128, h, b);
          }
        }
      for (int s = 0; s < nmax; ++s) {
        const int stage = s % Cfg::kStages;
        mbar_wait(&st_empty[stage], ((s / Cfg::kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&st_full[stage], Cfg::kStageBytes);
        unsigned char *dst = sStage + stage * Cfg::kStageBytes;
#pragma unroll
        for (int c = 0; c < Cfg::kChunks; ++c) {
          tma_load_4d(dst + c * Cfg::kChunk64, &tmK, &st_full[stage], c * 64, s * 64, h, b);
          tma_load_4d(dst + Cfg::kTile64 + c * Cfg::kChunk64, &tmV, &st_full[stage], c * 64, s * 64, h, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    if (elect_one()) {
      constexpr uint32_t idesc_xy = make_idesc(128, 64, IS_BF16, 0, 0);
      constexpr uint32_t idesc_acc = make_idesc(128, D, IS_BF16, 0, 1);
      const uint32_t sRes_a = smem_u32(sRes), sSt_a = smem_u32(sStage);
      auto wait_stage = [&](int s) {
        mbar_wait(&st_full[s % Cfg::kStages], (s / Cfg::kStages) & 1);
        tc_fence_after();
      };
      auto issue_xy = [&](int t, int s) {
        const uint32_t q_a = sRes_a + t * 2 * Cfg::kTile128, do_a = q_a + Cfg::kTile128;
        const uint32_t k_a = sSt_a + (s % Cfg::kStages) * Cfg::kStageBytes, v_a = k_a + Cfg::kTile64;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)  // S = Q K^T
          mma_ss(tmem_base + t * 128, kmajor_desc<Cfg::kChunk128>(q_a, kk),
                 kmajor_desc<Cfg::kChunk64>(k_a, kk), idesc_xy, kk > 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)  // dP = dO V^T
          mma_ss(tmem_base + t * 128 + 64, kmajor_desc<Cfg::kChunk128>(do_a, kk),
                 kmajor_desc<Cfg::kChunk64>(v_a, kk), idesc_xy, kk > 0);
        tc_commit(&xy_full[t]);
      };
      auto issue_acc = [&](int t, int s) {
        mbar_wait(&ds_full[t], s & 1);
        tc_fence_after();
        const uint32_t k_a = sSt_a + (s % Cfg::kStages) * Cfg::kStageBytes;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dQ += dS K   (K = 64 keys)
          mma_ts(tmem_base + 256 + t * D, tmem_base + t * 128 + 64 + kk * 8,
                 mnmajor_desc<Cfg::kChunk64>(k_a, kk), idesc_acc, (s > 0 || kk > 0) ? 1u : 0u);
        if (s == n_t[t] - 1) tc_commit(&acc_full[t]);
      };
      wait_stage(0);
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (n_t[t] > 0) {
          mbar_wait(&res_full[t], 0);
          tc_fence_after();
          issue_xy(t, 0);
        }
      for (int s = 0; s < nmax; ++s) {
        if (s + 1 < nmax) wait_stage(s + 1);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (s < n_t[t]) issue_acc(t, s);
          if (s + 1 < n_t[t]) issue_xy(t, s + 1);
        }
        tc_commit(&st_empty[s % Cfg::kStages]);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int D, int IS_BF16>
int launch_bwd_impl(const void *O, const void *dO, float *delta, const CUtensorMap *maps64,
                    const CUtensorMap *maps128, const BwdParams &p, int B, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    FA_CUDA_CHECK(cudaFuncSetAttribute(bwd_dkdv_kernel<D, IS_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       DkdvCfg<D>::kSmemBytes));
    FA_CUDA_CHECK(cudaFuncSetAttribute(bwd_dq_kernel<D, IS_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       DqCfg<D>::kSmemBytes));
    configured = true;
  }
  // maps: [0] Q, [1] K, [2] V, [3] dO
  bwd_delta_kernel<D, IS_BF16><<<dim3((p.N + 7) / 8, p.H, B), 256, 0, stream>>>(
      reinterpret_cast<const uint16_t *>(O), reinterpret_cast<const uint16_t *>(dO), delta, p.N, p.H,
      p.batch_stride, p.head_stride);
  FA_CUDA_CHECK(cudaGetLastError());
  bwd_dkdv_kernel<D, IS_BF16><<<dim3((p.N + 127) / 128, p.H, B), kBwdThreads, DkdvCfg<D>::kSmemBytes, stream>>>(
      maps64[0], maps128[1], maps128[2], maps64[3], p);
  FA_CUDA_CHECK(cudaGetLastError());
  bwd_dq_kernel<D, IS_BF16><<<dim3((p.N + 255) / 256, p.H, B), kBwdThreads, DqCfg<D>::kSmemBytes, stream>>>(
      maps128[0], maps64[1], maps64[2], maps128[3], p);
  FA_CUDA_CHECK(cudaGetLastError());
  count_launch(3);
  return FA_OK;
}

}  // namespace

int launch_bwd_tc(const void *Q, const void *K, const void *V, const void *O, const void *dO,
                  const float *L, float *dQ, float *dK, float *dV, int N, int D, float scale,
                  int64_t batch_stride, int64_t head_stride, int is_causal, int B, int H, int dtype,
                  void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  FA_REQUIRE(Q && K && V && O && dO && L && dQ && dK && dV, "null tensor pointer");
  FA_REQUIRE(N >= 1, "N must be >= 1 (got %d)", N);
  FA_REQUIRE(D == 64 || D == 128, "D must be 64 or 128 (got %d)", D);
  FA_REQUIRE(B >= 1 && H >= 1 && H <= 65535 && B <= 65535, "bad B/H (%d, %d)", B, H);
  FA_REQUIRE(dtype == FA_DTYPE_FP16 || dtype == FA_DTYPE_BF16, "dtype must be FA_DTYPE_FP16 or FA_DTYPE_BF16");
  FA_REQUIRE(scale > 0.f, "scale must be positive");
  FA_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(O) && aligned16(dO) && aligned16(dQ) &&
                 aligned16(dK) && aligned16(dV),
             "tensors must be 16-byte aligned");
  FA_REQUIRE(batch_stride % D == 0 && head_stride % D == 0, "strides must be multiples of D");
  FA_REQUIRE((H == 1 || head_stride >= (int64_t)N * D) && (B == 1 || batch_stride >= (int64_t)N * D),
             "heads overlap: stride smaller than N*D");
  // delta is indexed like L: offset / D + row; its extent follows the strides
  const size_t need = fa_workspace_bytes_backward(N, D, B, H);
  const int64_t last = ((int64_t)(B - 1) * batch_stride + (int64_t)(H - 1) * head_stride) / D + N;
  if (workspace == nullptr || workspace_bytes < need || (size_t)last * sizeof(float) > workspace_bytes)
    return set_error(FA_ERR_WORKSPACE,
                     "backward workspace too small: need %zu bytes (fa_workspace_bytes_backward; contiguous "
                     "[B,H,N,D] layout assumed), got %zu",
                     need, workspace_bytes);
  CUtensorMap m64[4], m128[4];
  const void *ptrs[4] = {Q, K, V, dO};
  int rc;
  for (int i = 0; i < 4; ++i) {
    if ((rc = make_tensor_map_bhnd(&m64[i], ptrs[i], dtype, N, D, H, B, head_stride, batch_stride, 64)) != FA_OK) return rc;
    if ((rc = make_tensor_map_bhnd(&m128[i], ptrs[i], dtype, N, D, H, B, head_stride, batch_stride, 128)) != FA_OK) return rc;
  }
  BwdParams p;
  p.L = L;
  p.delta = reinterpret_cast<const float *>(workspace);
  p.dQ = dQ; p.dK = dK; p.dV = dV;
  p.N = N; p.H = H;
  p.scale = scale;
  p.scale_log2 = scale * kLog2e;
  p.batch_stride = batch_stride;
  p.head_stride = head_stride;
  p.causal = is_causal ? 1 : 0;
  float *delta = reinterpret_cast<float *>(workspace);
  if (D == 64)
    return dtype == FA_DTYPE_BF16 ? launch_bwd_impl<64, 1>(O, dO, delta, m64, m128, p, B, stream)
                                  : launch_bwd_impl<64, 0>(O, dO, delta, m64, m128, p, B, stream);
  return dtype == FA_DTYPE_BF16 ? launch_bwd_impl<128, 1>(O, dO, delta, m64, m128, p, B, stream)
                                : launch_bwd_impl<128, 0>(O, dO, delta, m64, m128, p, B, stream);
}

}  // namespace fa
