// Fused 16-bit FlashAttention backward for sm_100a: ONE kernel, FIVE GEMMs per (key tile, query
// tile) pair -- S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q, dQ_i += dS K -- instead of
// the seven of the two-kernel form (bwd_tc.cu), and still no float atomics and bitwise run-to-run
// determinism (the reference accumulates with float atomics, kernels.metal:1227, 1243).
//
// A CTA owns one 128-key tile (K_j, V_j resident in shared memory, dK_j / dV_j accumulating in TMEM)
// and walks the 128-row query tiles from the LAST one down.  The part of dQ_i it produces is a fresh
// 128 x D product per pair; it leaves the SM through shared memory and a TMA add-reduction into the
// fp32 dQ tensor (cp.reduce.async.bulk.tensor .add -- performed by the L2, whole 128-byte lines).
// What makes that sum deterministic is its ORDER: the contributors of dQ tile i add strictly in
// ascending key-tile order, enforced with one counter per (head, query tile) in global memory
// (ld.acquire / red.release; zeroed before the launch).  Key tile 0 is always first and STORES
// (so dQ needs no zero fill) unless the caller asked to accumulate onto dQ (ring attention).
// CTAs are dispatched in ascending key-tile order and every CTA only ever waits for lower key tiles
// of its own head, so a waiting CTA's predecessor is always resident or finished: no deadlock.
// Walking the query tiles downwards makes all key tiles of a head meet a given dQ tile at the same
// iteration, so a CTA trails its predecessor by one reduction latency, once, instead of one
// iteration per query tile.
//
//   warps 0-3 / 4-7  element-wise warpgroups: thread = key row (TMEM lane), warpgroup w owns query
//                    columns [64w, 64w+64) of S^T / dP^T:  P^T = exp2(S^T c - L log2e) -> 16-bit
//                    back into TMEM (A operand of dV);  dS^T = P^T o (dP^T - D) scale -> 16-bit into
//                    TMEM (A operand of dK) AND into shared memory (operand of dQ, which needs
//                    dS with the QUERY as the M index);  then drains the previous pair's dQ block
//                    (thread = query row) into the swizzled staging boxes
//   warp 8           MMA issuer (one elected thread), tensor-pipe order per pair m:
//                       X(m+1) | dK(m) dQ(m) | dV(m+1) | Y(m+1)
//   warp 9           TMA producer: K_j, V_j once; Q_i / dO_i tiles through 2+1 (D=128) or 3+2 slots
//   warp 10          dQ reducer (one elected thread): ordering counter, TMA store / add-reduction
//   warp 11          idle (completes the warpgroup for setmaxnreg)
//
// TMEM, D = 128: X = S^T/P^T [0,128)  Y = dP^T/dS^T [128,256)  dV [256,384)  dK [384,512); the dQ
// block of pair m is written over Y once dS^T(m) has been consumed, and Y(m+1) waits until the
// element-wise threads have drained it.  D = 64: dV [256,320) dK [320,384) dQ [384,448).
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fa_internal.h"
#include "sched.cuh"
#include "sm100_ptx.cuh"
#include "tensormap.h"

namespace fa {
namespace {

using namespace ptx;

constexpr int kThreads = 384;
constexpr int kMmaWarp = 8;
constexpr int kLoadWarp = 9;
constexpr int kDqWarp = 10;
// 8 x 208 + 4 x 88 = 2016 = 12 x 168 registers per lane: the MMA issuer keeps its descriptors in registers
#ifndef FA_FUSED_REGS_WIDE
#define FA_FUSED_REGS_WIDE 208
#endif
constexpr int kRegsWide = FA_FUSED_REGS_WIDE, kRegsNarrow = (2016 - 8 * FA_FUSED_REGS_WIDE) / 4;
constexpr float kLog2e = 1.4426950408889634f;
#ifndef FA_BWD_FUSED_EMU
#define FA_BWD_FUSED_EMU 3  // every 3rd pair of exponentials on the FMA pipe (as in bwd_tc.cu)
#endif

__device__ __forceinline__ bool emulate_pair(int pair_index) {
  return FA_BWD_FUSED_EMU > 0 && (pair_index % (FA_BWD_FUSED_EMU > 0 ? FA_BWD_FUSED_EMU : 1)) == (FA_BWD_FUSED_EMU > 0 ? FA_BWD_FUSED_EMU : 1) - 1;
}
__device__ __forceinline__ uint64_t exp2_pair(uint64_t x2, bool emulate) {
  return emulate ? exp2_emulated_x2(x2) : pack_f32x2(ex2(lo_f32(x2)), ex2(hi_f32(x2)));
}

struct FusedParams {
  const float *L;      // [B, H, Nq] log-sum-exp of the scaled scores (natural log)
  const float *delta;  // [B, H, Nq] D_i = rowsum(O o dO)
  float *dK, *dV;
  uint32_t *sems;      // [B * H, n_q_tiles] ordering counters, zero at launch
  int Nq, Nk, H;
  float scale, scale_log2;
  int64_t batch_stride, head_stride;        // elements, of Q / dO / dQ (and L, delta via / D)
  int64_t kv_batch_stride, kv_head_stride;  // elements, of K / V / dK / dV
  int causal;          // requires Nq == Nk
  int acc_dq;          // dQ += (ring attention) instead of dQ =
  int group, n_heads;  // dispatch order (sched.cuh)
#ifdef FA_BWD_TRACE
  long long *prof;     // trace builds only: where one CTA's roles waited (cycles), see tests/fused_probe.py
#endif
};

// FA_BWD_TRACE (development builds only): the CTA of key tile 8 of head 0 accumulates the cycles each of
// its roles spends in every wait.  T_BEGIN(var) ... T_END(var, slot) bracket a region.
#ifdef FA_BWD_TRACE
#define T_DECL(on) const bool t_on = p.prof != nullptr && (on); long long t_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long t_0 = 0, t_start = t_on ? clock64() : 0
#define T_BEGIN() do { if (t_on) t_0 = clock64(); } while (0)
#define T_END(slot) do { if (t_on) t_acc[slot] += clock64() - t_0; } while (0)
#define T_FLUSH(base, count) do { if (t_on) { for (int t_i = 0; t_i < (count); ++t_i) p.prof[(base) + t_i] = t_acc[t_i]; p.prof[(base) + (count)] = clock64() - t_start; } } while (0)
#else
#define T_DECL(on) do {} while (0)
#define T_BEGIN() do {} while (0)
#define T_END(slot) do {} while (0)
#define T_FLUSH(base, count) do {} while (0)
#endif

template <int D>
struct FusedCfg {
  static constexpr int kChunk = 128 * 128;              // [128 rows][64 elements] 16-bit
  static constexpr int kTile = (D / 64) * kChunk;       // one 128-row Q / K / V / dO tile
  static constexpr int kQSlots = D == 128 ? 2 : 3;
  static constexpr int kDoSlots = D == 128 ? 1 : 2;
  static constexpr int kDsBytes = 2 * kChunk;           // dS^T [128 keys][128 queries] 16-bit
  static constexpr int kBoxBytes = 128 * 128;           // dQ staging box [128 rows][32 fp32]
  static constexpr int kBoxesPerWg = D / 64;            // boxes a warpgroup fills per pair (one buffer each)
  static constexpr int kSmemTiles = (2 + kQSlots + kDoSlots) * kTile + kDsBytes + 2 * kBoxBytes;
  static constexpr int kStatBytes = 2 * 2 * 128 * 4;    // [wg][double buffer][-L log2e 64 | -D scale 64]
  static constexpr int kSmemBytes = kSmemTiles + kStatBytes + 512 + 256;
  static constexpr int kTmemDV = 256, kTmemDK = 256 + D;
  static constexpr int kTmemDQ = D == 128 ? 128 : 384;  // D = 128: over Y
};

__device__ __forceinline__ uint64_t kmajor_desc(uint32_t tile_addr, int kk) {
  return make_sdesc_sw128(tile_addr + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t tile_addr, int kk) {
  return make_sdesc_sw128(tile_addr + kk * 2048, 128 * 128, 1024);
}
template <int D>
__device__ __forceinline__ void tma_load_tile(unsigned char *dst, const CUtensorMap *map, uint64_t *bar, int row, int h, int b) {
#pragma unroll
  for (int c = 0; c < D / 64; ++c) tma_load_4d(dst + c * (128 * 128), map, bar, c * 64, row, h, b);
}

template <int D, int IS_BF16>
__global__ void __launch_bounds__(kThreads, 1)
bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                 const __grid_constant__ CUtensorMap tmdQ, const FusedParams p) {
  using Cfg = FusedCfg<D>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  if (smem - smem_raw > 512) __trap();
  unsigned char *sK = smem;
  unsigned char *sV = smem + Cfg::kTile;
  unsigned char *sQ = smem + 2 * Cfg::kTile;                      // [kQSlots]
  unsigned char *sDO = sQ + Cfg::kQSlots * Cfg::kTile;            // [kDoSlots]
  unsigned char *sDS = sDO + Cfg::kDoSlots * Cfg::kTile;          // dS^T, K-major [key][query], 2 chunks
  unsigned char *sBox = sDS + Cfg::kDsBytes;                      // [2 warpgroups] dQ staging
  float *sLD = reinterpret_cast<float *>(smem + Cfg::kSmemTiles);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemTiles + Cfg::kStatBytes);
  uint64_t *res_full = bars;          // K_j, V_j landed
  uint64_t *acc_full = bars + 1;      // dV, dK final
  uint64_t *x_full = bars + 2;        // S^T ready
  uint64_t *y_full = bars + 3;        // dP^T ready
  uint64_t *dq_full = bars + 4;       // dQ block of this pair ready in TMEM
  uint64_t *dq_drained = bars + 5;    // ... copied to registers by every element-wise thread
  uint64_t *p_ready = bars + 6;       // [2] P^T part stored by both warpgroups
  uint64_t *ds_ready = bars + 8;      // [2] dS^T part stored (TMEM and shared memory) by both warpgroups
  uint64_t *stage_full = bars + 10;   // [2] warpgroup w has filled its staging box
  uint64_t *stage_free = bars + 12;   // [2] the TMA engine has read it
  uint64_t *q_full = bars + 14;                      // [kQSlots]
  uint64_t *q_empty = q_full + Cfg::kQSlots;
  uint64_t *do_full = q_empty + Cfg::kQSlots;        // [kDoSlots]
  uint64_t *do_empty = do_full + Cfg::kDoSlots;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(do_empty + Cfg::kDoSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // dispatch order (sched.cuh): key tile 0 first, ascending -- also the order of the dQ reduction
  const BlockCoord bc = decode_block(p.group, p.n_heads, p.H);
  if (bc.b < 0) return;
  const int j = bc.blk, h = bc.h, b = bc.b;
  const int key0 = j * 128;
  const int n_q_tiles = (p.Nq + 127) / 128;
  const int t_lo = p.causal ? j : 0;       // first query tile that sees these keys
  const int n = n_q_tiles - t_lo;          // pairs of this CTA, query tile of pair m: n_q_tiles - 1 - m
  const int64_t kv_off = (int64_t)b * p.kv_batch_stride + (int64_t)h * p.kv_head_stride;
  const int64_t vec_off = ((int64_t)b * p.batch_stride + (int64_t)h * p.head_stride) / D;
#ifdef FA_BWD_TRACE
  const bool traced_cta = j == 8 && h == 0 && b == 0;
#endif

  if (threadIdx.x == 0) {
    mbar_init(res_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(x_full, 1);
    mbar_init(y_full, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_drained, 8 * kArrivalsPerWarp);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_ready[i], 8 * kArrivalsPerWarp);
      mbar_init(&ds_ready[i], 8 * kArrivalsPerWarp);
      mbar_init(&stage_full[i], 4 * kArrivalsPerWarp);
      mbar_init(&stage_free[i], 1);
    }
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
    setmaxnreg_inc<kRegsWide>();
    const int wg = warp >> 2;
    const int tid = (warp & 3) * 32 + lane;  // TMEM lane: key row (S^T, dP^T) / query row (dQ block)
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tX = tmem_base + lane_off + wg * 64;         // this warpgroup's half of S^T / P^T
    const uint32_t tY = tmem_base + lane_off + 128 + wg * 64;   // ... of dP^T / dS^T
    const uint32_t tDQ = tmem_base + lane_off + Cfg::kTmemDQ + wg * (D / 2);  // its columns of the dQ block
    const int key = key0 + tid;
    // per-column statistics (-L_i log2e for threads 0-63, -D_i scale for threads 64-127) of the 64 query
    // columns this warpgroup owns, fetched one pair ahead (see bwd_tc.cu: the value is only touched at
    // the top of the next iteration so the global load latency is never exposed)
    const float *stat_src = (tid < 64 ? p.L : p.delta) + vec_off + wg * 64 + (tid & 63);
    const float stat_coef = tid < 64 ? -kLog2e : -p.scale;
    auto fetch_stat = [&](int m) -> float {
      const int q_first = (n_q_tiles - 1 - m) * 128;
      float v = tid < 64 ? CUDART_INF_F : 0.f;  // rows past N: P = exp2(-inf) = 0, D = 0
      if (m < n && q_first + wg * 64 + (tid & 63) < p.Nq) v = __ldg(stat_src + q_first);
      return v;
    };
    float stat_next = fetch_stat(0);
    const uint64_t scale_log2_2 = pack_f32x2(p.scale_log2, p.scale_log2), scale_2 = pack_f32x2(p.scale, p.scale);
    // my 128-byte row of the shared-memory dS^T chunk of this warpgroup (128-byte swizzle: 16-byte unit u
    // of row r lives at unit u ^ (r & 7)) and of the staging box
    const uint32_t ds_row = smem_u32(sDS) + wg * Cfg::kChunk + tid * 128;
    const uint32_t box_row = smem_u32(sBox) + wg * Cfg::kBoxBytes + tid * 128;
    const uint32_t sw = (uint32_t)(tid & 7);
    T_DECL(traced_cta && tid == 0 && wg == 0);  // slots: 0 x_full 1 phase1 2 dq_full 3 stage_free(A) 4 y_full 5 phase2 6 stage_free(B) 7 stats barrier
    int fills = 0;  // staging-box fills of this warpgroup so far
    // dQ block of a pair: TMEM -> registers (frees the TMEM block) -> staging box -> reducer warp.  The
    // warpgroup has ONE 16 KB box; at D = 128 it fills it twice per pair.  The second fill has to wait
    // until the TMA engine has read the first, so it is deferred until after phase 2 of the current pair
    // (the values wait in registers) instead of stalling here.
    uint32_t dqv[Cfg::kBoxesPerWg][32];
    auto fill_box = [&](int hb) {
      T_BEGIN();
      if (fills > 0) mbar_wait(&stage_free[wg], (fills - 1) & 1);
      T_END(hb == 0 ? 3 : 6);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        sts_v4(box_row + ((u ^ sw) << 4), dqv[hb][4 * u], dqv[hb][4 * u + 1], dqv[hb][4 * u + 2], dqv[hb][4 * u + 3]);
      fence_proxy_async();
      mbar_arrive_warp(&stage_full[wg]);
      ++fills;
    };
    auto drain_dq = [&](int m) {
      T_BEGIN();
      mbar_wait(dq_full, m & 1);
      T_END(2);
      tc_fence_after();
#pragma unroll
      for (int hb = 0; hb < Cfg::kBoxesPerWg; ++hb) tmem_ld32(tDQ + hb * 32, dqv[hb]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive_warp(dq_drained);
      fill_box(0);
    };
    for (int m = 0; m < n; ++m) {
      const int q0 = (n_q_tiles - 1 - m) * 128 + wg * 64;  // first query column of this warpgroup's half
      float *ld = sLD + (wg * 2 + (m & 1)) * 128;
      T_BEGIN();
      ld[tid] = stat_next * stat_coef;
      stat_next = fetch_stat(m + 1);
      named_bar_sync(1 + wg, 128);
      T_END(7);
      // ---- phase 1: P^T = exp2(S^T * c - L * log2e) ----
      T_BEGIN();
      mbar_wait(x_full, m & 1);
      T_END(0);
      T_BEGIN();
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
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive_warp(&p_ready[c]);
      }
      T_END(1);
      // ---- the previous pair's dQ block leaves TMEM (Y(m) is only issued after this drain) ----
      if (m > 0) drain_dq(m - 1);
      // ---- phase 2: dS^T = P^T o (dP^T * scale - D * scale) ----
      T_BEGIN();
      mbar_wait(y_full, m & 1);
      T_END(4);
      T_BEGIN();
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
        // the same 32 query columns (four 16-byte units 4c .. 4c+3 of my row) into the shared-memory copy
#pragma unroll
        for (int u = 0; u < 4; ++u)
          sts_v4(ds_row + (((uint32_t)(4 * c + u) ^ sw) << 4), dk[4 * u], dk[4 * u + 1], dk[4 * u + 2], dk[4 * u + 3]);
        fence_proxy_async();
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive_warp(&ds_ready[c]);
      }
      T_END(5);
      if (Cfg::kBoxesPerWg == 2 && m > 0) fill_box(1);  // second half of the previous pair's dQ block
    }
    drain_dq(n - 1);
    if (Cfg::kBoxesPerWg == 2) fill_box(1);
    T_FLUSH(0, 8);
#ifdef FA_BWD_TRACE
    if (t_on) p.prof[63] = n;
#endif
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
    setmaxnreg_dec<kRegsNarrow>();
    if (warp == kLoadWarp) {
      // ============================ TMA producer ============================
      if (elect_one()) {
        prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmdO);
        mbar_arrive_expect_tx(res_full, 2 * Cfg::kTile);
        tma_load_tile<D>(sK, &tmK, res_full, key0, h, b);
        tma_load_tile<D>(sV, &tmV, res_full, key0, h, b);
        for (int m = 0; m < n; ++m) {
          const int q0 = (n_q_tiles - 1 - m) * 128;
          const int qs = m % Cfg::kQSlots, ds = m % Cfg::kDoSlots;
          mbar_wait(&q_empty[qs], ((m / Cfg::kQSlots) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[qs], Cfg::kTile);
          tma_load_tile<D>(sQ + qs * Cfg::kTile, &tmQ, &q_full[qs], q0, h, b);
          mbar_wait(&do_empty[ds], ((m / Cfg::kDoSlots) & 1) ^ 1);
          mbar_arrive_expect_tx(&do_full[ds], Cfg::kTile);
          tma_load_tile<D>(sDO + ds * Cfg::kTile, &tmdO, &do_full[ds], q0, h, b);
        }
      }
      __syncwarp();
    } else if (warp == kMmaWarp) {
      // ============================= MMA issuer =============================
      if (elect_one()) {
        constexpr uint32_t idesc_xy = make_idesc(128, 128, IS_BF16, 0, 0);   // S^T, dP^T: both operands K-major
        constexpr uint32_t idesc_acc = make_idesc(128, D, IS_BF16, 0, 1);    // dV, dK: A from TMEM, B MN-major
        constexpr uint32_t idesc_dq = make_idesc(128, D, IS_BF16, 1, 1);     // dQ: A = dS (MN-major view of dS^T), B MN-major
        const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV), sQ_a = smem_u32(sQ), sDO_a = smem_u32(sDO), sDS_a = smem_u32(sDS);
        const uint32_t tX = tmem_base, tY = tmem_base + 128, tdV = tmem_base + Cfg::kTmemDV, tdK = tmem_base + Cfg::kTmemDK,
                       tdQ = tmem_base + Cfg::kTmemDQ;
        T_DECL(traced_cta);  // slots: 0 q_full 1 do_full 2 p_ready 3 ds_ready 4 dq_drained
        auto q_addr = [&](int m) { return sQ_a + (m % Cfg::kQSlots) * Cfg::kTile; };
        auto do_addr = [&](int m) { return sDO_a + (m % Cfg::kDoSlots) * Cfg::kTile; };
        auto issue_x = [&](int m) {  // S^T = K Q^T
          T_BEGIN();
          mbar_wait(&q_full[m % Cfg::kQSlots], (m / Cfg::kQSlots) & 1);
          T_END(0);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)
            mma_ss(tX, kmajor_desc(sK_a, kk), kmajor_desc(q_addr(m), kk), idesc_xy, kk > 0);
          tc_commit(x_full);
        };
        auto issue_y = [&](int m) {  // dP^T = V dO^T
          mbar_wait(&do_full[m % Cfg::kDoSlots], (m / Cfg::kDoSlots) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)
            mma_ss(tY, kmajor_desc(sV_a, kk), kmajor_desc(do_addr(m), kk), idesc_xy, kk > 0);
          tc_commit(y_full);
        };
        auto issue_dv = [&](int m) {  // dV += P^T dO (K = 128 query rows); part c = k-steps {2c, 2c+1, 4+2c, 5+2c}
          T_BEGIN();
          mbar_wait(&do_full[m % Cfg::kDoSlots], (m / Cfg::kDoSlots) & 1);
          T_END(1);
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            T_BEGIN();
            mbar_wait(&p_ready[part], m & 1);
            T_END(2);
            tc_fence_after();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int kk = (q >> 1) * 4 + part * 2 + (q & 1);
              mma_ts(tdV, tX + (kk >> 2) * 64 + (kk & 3) * 8, mnmajor_desc(do_addr(m), kk), idesc_acc,
                     (m > 0 || part > 0 || q > 0) ? 1u : 0u);
            }
          }
        };
        mbar_wait(res_full, 0);
        tc_fence_after();
        issue_x(0);
        issue_y(0);
        issue_dv(0);
        if (Cfg::kDoSlots == 1 || n > 1) tc_commit(&do_empty[0]);
        for (int m = 0; m < n; ++m) {
          if (m + 1 < n) issue_x(m + 1);  // X is free: dV(m) was issued before
          // dK += dS^T Q (A from TMEM), then the dQ block = dS K (both operands from shared memory)
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            T_BEGIN();
            mbar_wait(&ds_ready[part], m & 1);
            T_END(3);
            tc_fence_after();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int kk = (q >> 1) * 4 + part * 2 + (q & 1);
              mma_ts(tdK, tY + (kk >> 2) * 64 + (kk & 3) * 8, mnmajor_desc(q_addr(m), kk), idesc_acc,
                     (m > 0 || part > 0 || q > 0) ? 1u : 0u);
            }
          }
          tc_commit(&q_empty[m % Cfg::kQSlots]);
          if (m == n - 1) tc_commit(acc_full);
          // k-step kk = key rows 16kk .. 16kk+15 of dS^T: the whole tile has been stored (both parts)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            mma_ss(tdQ, mnmajor_desc(sDS_a, kk), mnmajor_desc(sK_a, kk), idesc_dq, kk > 0);
          tc_commit(dq_full);
          if (m + 1 < n) {
            issue_dv(m + 1);
            T_BEGIN();
            mbar_wait(dq_drained, m & 1);  // the dQ block (over Y at D = 128) has left TMEM
            T_END(4);
            tc_fence_after();
            issue_y(m + 1);
            tc_commit(&do_empty[(m + 1) % Cfg::kDoSlots]);
          }
        }
        T_FLUSH(16, 5);
      }
      __syncwarp();
    } else if (warp == kDqWarp) {
      // ============================== dQ reducer ==============================
      if (elect_one()) {
        prefetch_tensormap(&tmdQ);
        uint32_t *sem = p.sems + (size_t)(b * p.H + h) * n_q_tiles;
        const bool store_first = (j == 0) && !p.acc_dq;  // key tile 0 is the first contributor of every dQ tile
        T_DECL(traced_cta);  // slots: 0 stage_full 1 ordering counter 2 read wait 3 completion wait
        int fills = 0;  // fills consumed per warpgroup (both advance together)
        for (int m = 0; m < n; ++m) {
          const int ti = n_q_tiles - 1 - m;
#pragma unroll
          for (int hb = 0; hb < Cfg::kBoxesPerWg; ++hb) {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              T_BEGIN();
              mbar_wait(&stage_full[w], fills & 1);
              T_END(0);
              T_BEGIN();
              if (hb == 0 && w == 0 && j > 0) {
                // my turn: key tiles 0 .. j-1 have finished adding their part of dQ tile ti
                const long long t0 = clock64();
                while (ld_acquire_gpu(sem + ti) < (uint32_t)j) {
                  __nanosleep(100);
                  if (clock64() - t0 > 4000000000LL) __trap();
                }
              }
              T_END(1);
              const unsigned char *box = sBox + w * Cfg::kBoxBytes;
              const int col = w * (D / 2) + hb * 32;
              if (store_first) tma_store_4d(&tmdQ, box, col, ti * 128, h, b);
              else tma_reduce_add_4d(&tmdQ, box, col, ti * 128, h, b);
              bulk_commit_group();
            }
            T_BEGIN();
            bulk_wait_read_all();  // both boxes have been read: the warpgroups may refill them
            T_END(2);
            mbar_arrive(&stage_free[0]);
            mbar_arrive(&stage_free[1]);
            ++fills;
          }
          // The groups of pair m are in flight; those of pair m - 1 are complete once at most this pair's
          // groups are pending: only then may the next key tile add to dQ tile ti + 1 (release, one pair late,
          // instead of idling here until the L2 has performed this pair's reduction).
          if (m > 0) {
            T_BEGIN();
            bulk_wait_pending<2 * Cfg::kBoxesPerWg>();
            T_END(3);
            __threadfence();
            red_release_gpu_add(sem + ti + 1, 1u);
          }
        }
        bulk_wait_all();
        __threadfence();
        red_release_gpu_add(sem + (n_q_tiles - n), 1u);  // the last pair's query tile
        T_FLUSH(32, 4);
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int D, int IS_BF16>
int configure() {
  static DeviceOnce configured;
  return configured.run([] {
    FA_CUDA_CHECK(cudaFuncSetAttribute(bwd_fused_kernel<D, IS_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       FusedCfg<D>::kSmemBytes));
    cudaFuncAttributes attr;  // forces the (lazily loaded) kernel into the context
    FA_CUDA_CHECK(cudaFuncGetAttributes(&attr, bwd_fused_kernel<D, IS_BF16>));
    return (int)FA_OK;
  });
}

template <int D, int IS_BF16>
int launch_impl(const CUtensorMap *const *maps, const FusedParams &p, int B, cudaStream_t stream) {
  const int rc = configure<D, IS_BF16>();
  if (rc != FA_OK) return rc;
  FusedParams q = p;
  q.n_heads = B * p.H;
  const int n_k_tiles = (p.Nk + 127) / 128;
  // heads interleaved in L2-sized groups for every launch (not only causal ones): it keeps the chain of
  // CTAs that wait for each other on one head's dQ tiles short
  q.group = dispatch_group(true, (int64_t)2 * (p.Nq > p.Nk ? p.Nq : p.Nk) * D * 2, q.n_heads);
  if (n_k_tiles > 65535) q.group = 1;  // grid.y limit of the grouped form
  bwd_fused_kernel<D, IS_BF16><<<dispatch_grid(q.group, n_k_tiles, p.H, B), kThreads, FusedCfg<D>::kSmemBytes, stream>>>(
      *maps[0], *maps[1], *maps[2], *maps[3], *maps[4], q);
  FA_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return FA_OK;
}

}  // namespace

int preload_bwd_fused() {
  int rc;
  if ((rc = configure<64, 0>()) || (rc = configure<64, 1>()) || (rc = configure<128, 0>()) || (rc = configure<128, 1>())) return rc;
  return FA_OK;
}

size_t bwd_fused_sem_bytes(int Nq, int B, int H) {
  const size_t n = (size_t)B * H * ((Nq + 127) / 128) * sizeof(uint32_t);
  return (n + 255) & ~(size_t)255;
}

// Fused backward of a (rectangular) block; arguments as launch_bwd_tc_rect (validated there).  `sems`
// holds bwd_fused_sem_bytes() bytes; it is zeroed here, in stream order.
int launch_bwd_fused(const void *Q, const void *K, const void *V, const void *dO, const float *L, const float *delta,
                     float *dQ, float *dK, float *dV, int Nq, int Nk, int D, float scale, int64_t q_batch_stride,
                     int64_t q_head_stride, int64_t kv_batch_stride, int64_t kv_head_stride, int is_causal, int acc_dq,
                     int B, int H, int dtype, void *sems, cudaStream_t stream) {
  const CUtensorMap *maps[5];
  int rc;
  if ((rc = tensor_map_bhnd(&maps[0], Q, dtype, Nq, D, H, B, q_head_stride, q_batch_stride, 128)) != FA_OK) return rc;
  if ((rc = tensor_map_bhnd(&maps[1], K, dtype, Nk, D, H, B, kv_head_stride, kv_batch_stride, 128)) != FA_OK) return rc;
  if ((rc = tensor_map_bhnd(&maps[2], V, dtype, Nk, D, H, B, kv_head_stride, kv_batch_stride, 128)) != FA_OK) return rc;
  if ((rc = tensor_map_bhnd(&maps[3], dO, dtype, Nq, D, H, B, q_head_stride, q_batch_stride, 128)) != FA_OK) return rc;
  if ((rc = tensor_map_bhnd(&maps[4], dQ, kTensorMapF32, Nq, D, H, B, q_head_stride, q_batch_stride, 128)) != FA_OK) return rc;
  FA_CUDA_CHECK(cudaMemsetAsync(sems, 0, bwd_fused_sem_bytes(Nq, B, H), stream));
  FusedParams p = {};
  p.L = L;
  p.delta = delta;
  p.dK = dK; p.dV = dV;
  p.sems = reinterpret_cast<uint32_t *>(sems);
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
    return dtype == FA_DTYPE_BF16 ? launch_impl<64, 1>(maps, p, B, stream) : launch_impl<64, 0>(maps, p, B, stream);
  return dtype == FA_DTYPE_BF16 ? launch_impl<128, 1>(maps, p, B, stream) : launch_impl<128, 0>(maps, p, B, stream);
}

}  // namespace fa
