// 16-bit FlashAttention forward for sm_100a: tcgen05 MMAs with TMEM accumulators,
// TMA-fed mbarrier pipeline, warp-specialised roles.
//
// Replaces flash_attention_v4_half_kernel (kernels.metal:600-883) and, called with
// B = H = 1 / L = NULL / causal = 0, flash_attention_simd_kernel (kernels.metal:177-455).
//
// One CTA per SM works on 256 query rows of one (batch, head): two 128-row Q tiles
// that share every K/V tile brought in by TMA (halves L2->SM traffic) and ping-pong
// on the tensor core so the softmax of one tile overlaps the MMAs of the other.
//
//   warps 0-3   softmax, Q tile 0   one thread per query row (TMEM lane = row):
//   warps 4-7   softmax, Q tile 1   tcgen05.ld S -> mask -> online softmax (exp2,
//                                   conditional rescale) -> tcgen05.st P (16-bit)
//                                   over S; epilogue O / l -> global, L
//   warp  8     MMA issuer          one elected thread: S_t = Q_t K_j^T (smem x smem),
//                                   O_t += P_t V_j (TMEM x smem), tcgen05.commit -> mbarriers
//   warp  9     TMA producer        one elected thread: Q tiles once, then K_j, V_j into a
//                                   ring of shared-memory stages
//   warps 10-11 idle (complete the warpgroup for setmaxnreg)
//
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [256+D,256+2D);
// P_t aliases the first 64 columns of S_t (two 16-bit values per column).
//
// P is published to the MMA warp in parts (4 at D = 128, 2 at D = 64), one mbarrier each; one pair
// of exponentials in four is computed on the FMA pipe (see the knobs below).
//
// Launch geometry (sched.cuh): heads in L2-sized groups, heaviest row blocks first.  Launches too
// small to fill the GPU run as thread-block clusters of 2 or 4 CTAs per row block, each CTA on its
// share of the key tiles, merged through distributed shared memory in the epilogue.
//
// Shared memory tiles are [128 rows][64 elements] boxes (128-byte rows, 128-byte
// swizzle) exactly as TMA writes them: K-major operands for Q K^T (Q and K rows are
// the M/N index, head dim is K), MN-major B operand for P V (V rows are the K index).
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <atomic>

#include "fa_internal.h"
#include "sched.cuh"
#include "sm100_ptx.cuh"
#include "tensormap.h"

namespace fa {
namespace {

using namespace ptx;

constexpr int kBM = 128;          // query rows per tile (= TMEM lanes)
constexpr int kBN = 128;          // keys per tile
constexpr int kThreads = 384;
constexpr int kMmaWarp = 8;
constexpr int kLoadWarp = 9;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // in log2 units: P may grow to 2^8 before O is rescaled
// Tuning knobs, measured on B200 (tests/perf_probe.py d64, flagship shapes):
//  * every FA_FWD_EMU-th pair of exponentials is computed on the FMA pipe instead of MUFU.EX2
//    (16 / clk / SM: 128 exponentials per row need as many MUFU cycles as the tile's MMAs need
//    tensor cycles at D = 128, twice as many at D = 64).  1 in 4 is the optimum for both head
//    dims: an emulated pair costs ~22 issue cycles against 8 + 16 MUFU-pipe cycles
//    (tools/pipe_rate_probe.cu), so a larger share makes the loop issue-bound.
//  * P is handed to the MMA warp in FA_FWD_PARTS pieces per tile (4 at D = 128: 32 keys = two
//    k-steps of O += P V each; 2 at D = 64 where the PV MMA is half as long).
// 0/2 -> 4/4 at D = 128: 1190 -> 1300 TFLOP/s causal, 1355 -> 1470 non-causal (N = 16384, H = 16);
// 0/1 -> 4/2 at D = 64: 637 -> 751 causal.
#ifndef FA_FWD_EMU64
#define FA_FWD_EMU64 4
#endif
#ifndef FA_FWD_EMU128
#define FA_FWD_EMU128 4
#endif
#ifndef FA_FWD_PARTS64
#define FA_FWD_PARTS64 2
#endif
#ifndef FA_FWD_PARTS128
#define FA_FWD_PARTS128 4
#endif

template <int D>
struct FwdCfg {
  static constexpr int kChunks = D / 64;                    // 64-element (128 B) column chunks
  static constexpr int kChunkBytes = 128 * 128;             // 128 rows x 128 B
  static constexpr int kTileBytes = kChunks * kChunkBytes;  // one Q, K or V tile
  static constexpr int kStages = D == 128 ? 4 : 8;
  static constexpr int kSmemTiles = (2 + kStages) * kTileBytes;
  static constexpr int kSmemBytes = kSmemTiles + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// FA_FWD_TRACE (development builds only): one CTA records clock64() timestamps of its hand-offs
// for four consecutive iterations into the buffer set with fa_debug_set_prof_buffer.
#ifdef FA_FWD_TRACE
#define FA_TRACE(cond, j, slot)                                                                   \
  do {                                                                                            \
    if (p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (cond) &&   \
        (j) >= 8 && (j) < 12)                                                                     \
      p.prof[((j) - 8) * 64 + (slot)] = clock64();                                                \
  } while (0)
#else
#define FA_TRACE(cond, j, slot) do {} while (0)
#endif

struct FwdParams {
#ifdef FA_FWD_TRACE
  long long *prof;  // hand-off timestamps (trace builds only)
#endif
  void *O;
  float *L;
  int Nq;             // query rows (= rows of O and L)
  int Nk;             // keys (== Nq unless called for a rectangular ring-attention block)
  int H;
  float scale;        // multiplies the dot product
  float scale_log2;   // scale * log2(e)
  int64_t batch_stride, head_stride;  // elements, of Q / O
  int causal;         // requires Nq == Nk
  int group, n_blocks, n_heads;  // dispatch order (sched.cuh)
  int split;          // CTAs per cluster that share one row block, each taking 1/split of the keys
  // Ring attention: the launch is one (Q block x K/V chunk) partial of a longer softmax row.  The
  // epilogue folds it into a running fp32 (O_acc, L_acc) pair with the online-softmax rule
  // (kernels.metal:784-791) instead of a separate merge kernel; rows below `half_rows` (the first
  // zig-zag chunk) and the rest can be at different points of their sequence of partials.
  float *O_acc;       // fp32, addressed like O
  float *L_acc;       // fp32, addressed like L
  int merge_lo, merge_hi, half_rows;  // kMerge* per row range
};

enum { kMergeNone = 0, kMergeFirst = 1, kMergeMiddle = 2, kMergeLast = 3 };

template <int D, int IS_BF16>
__global__ void __launch_bounds__(kThreads, 1)
fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  using Cfg = FwdCfg<D>;
  extern __shared__ unsigned char smem_raw[];
  // 128-byte swizzle atoms are 1024 B: align the tile area
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char *sQ = smem;                           // 2 tiles
  unsigned char *sKV = smem + 2 * Cfg::kTileBytes;    // kStages tiles
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemTiles);
  uint64_t *q_full = bars;                       // [2]
  uint64_t *s_full = bars + 2;                   // [2]
  uint64_t *p_full = bars + 4;                   // [2 tiles][up to 4 parts of the key columns]
  uint64_t *o_full = bars + 12;                  // [2]
  uint64_t *kv_full = bars + 14;                 // [kStages]
  uint64_t *kv_empty = bars + 14 + Cfg::kStages; // [kStages]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 14 + 2 * Cfg::kStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // dispatch order: see sched.cuh (causal: heaviest = latest row blocks first, over groups of heads)
  const BlockCoord bc = decode_block(p.group, p.n_heads, p.H);
  if (bc.b < 0) return;
  const int h = bc.h, b = bc.b;
  // causal: the last row blocks have the most keys -> schedule them first
  // Small launches (fewer row blocks than half the SMs) run split = 2, 4 or 8: a cluster of CTAs
  // works on one row block, each on its share of the key tiles; the partial (m, l, O) results are
  // merged through distributed shared memory by the online-softmax rule (see the epilogue).
  const int split = p.split;
  const int crank = split > 1 ? (int)cluster_ctarank() : 0;
  const int blk = split > 1 ? bc.blk / split : bc.blk;
  const int qb = p.causal ? (p.n_blocks - 1 - blk) : blk;
  const int q_row0 = qb * 2 * kBM;
  const int n_kv_all = (p.Nk + kBN - 1) / kBN;
  // KV tiles each Q tile needs (0 = tile entirely past N)
  int n_t[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int r0 = q_row0 + t * kBM;
    n_t[t] = r0 >= p.Nq ? 0 : (p.causal ? min(n_kv_all, r0 / kBN + 1) : n_kv_all);
  }
  // this CTA's share of the key tiles: global tiles [j_begin, j_begin + c_t[t]) of Q tile t
  const int n_all = max(n_t[0], n_t[1]);
  const int j_begin = crank * n_all / split, j_end = (crank + 1) * n_all / split;
  int c_t[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) c_t[t] = max(0, min(j_end, n_t[t]) - j_begin);
  const int nmax = max(c_t[0], c_t[1]);
#ifdef FA_FWD_TRACE
  // SM clock under load: cycles and nanoseconds over the life of the last CTA in launch order
  if (p.prof != nullptr && threadIdx.x == 0 && blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.prof[256] = clock64();
    p.prof[257] = (long long)ns;
  }
#endif

#ifdef FA_FWD_TRACE
  if (p.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    p.prof[260] = clock64();
    p.prof[264] = nmax;
  }
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      for (int q = 0; q < 4; ++q) mbar_init(&p_full[4 * i + q], 4 * kArrivalsPerWarp);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // =========================== softmax warpgroups ===========================
    setmaxnreg_inc<216>();
    const int t = warp >> 2;
    const int row_in_tile = (warp & 3) * 32 + lane;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * kBN;
    const uint32_t tO = tmem_base + lane_off + 256 + t * D;
    const int grow = q_row0 + t * kBM + row_in_tile;
    const int nt = c_t[t];
    float m_run = -CUDART_INF_F;  // reference max (raw score units) the accumulators are relative to
    float l_run = 0.f;

    for (int j = 0; j < nt; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
#ifdef FA_FWD_TRACE
      if (p.prof != nullptr && threadIdx.x == 0 && j == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.prof[261] = clock64();
#endif
      FA_TRACE((warp & 3) == 0 && lane == 0, j, t * 5 + 0);
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, s[c]);
      tmem_wait_ld();
      FA_TRACE((warp & 3) == 0 && lane == 0, j, t * 5 + 1);

      // ---- masks: causal diagonal tile (always the last one) / keys past N ----
      const int gj = j_begin + j;  // key tile index in the sequence
      const bool diag = p.causal && (gj == n_t[t] - 1);
      const bool tail = (gj + 1) * kBN > p.Nk;
      if (diag || tail) {
        int limit = p.Nk - 1 - gj * kBN;                 // last valid key column in this tile
        if (diag) limit = min(limit, grow - gj * kBN);  // key <= query row
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i > limit) s[c][i] = 0xff800000u;  // -inf
      }
      // P = exp2(s * scale_log2 - m * scale_log2) for 32 keys of the row: packed 16-bit values
      // into 16 columns of S, partial row sums into sum2
      uint64_t sum2[2] = {0ull, 0ull};
      constexpr int kEmu = D == 64 ? FA_FWD_EMU64 : FA_FWD_EMU128;
      constexpr int kParts = D == 64 ? FA_FWD_PARTS64 : FA_FWD_PARTS128;  // P hand-offs per tile
      auto exp_chunk = [&](int c, float m_ref) {
        const float neg_m = -m_ref * p.scale_log2;
        // packed fp32x2 FMA / ADD: one issue slot per two elements (FFMA2 / FADD2)
        const uint64_t scale2 = pack_f32x2(p.scale_log2, p.scale_log2), negm2 = pack_f32x2(neg_m, neg_m);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t x2 = fma_f32x2(pack_u32x2(s[c][i], s[c][i + 1]), scale2, negm2);
          float p0, p1;
          if (kEmu > 0 && ((i >> 1) % (kEmu > 0 ? kEmu : 1)) == (kEmu > 0 ? kEmu : 1) - 1) {
            const uint64_t e2 = exp2_emulated_x2(x2);  // FMA pipe instead of MUFU
            p0 = lo_f32(e2);
            p1 = hi_f32(e2);
          } else {
            p0 = ex2(lo_f32(x2));
            p1 = ex2(hi_f32(x2));
          }
          sum2[(i >> 1) & 1] = add_f32x2(sum2[(i >> 1) & 1], pack_f32x2(p0, p1));
          pk[i >> 1] = pack2<IS_BF16>(p0, p1);
        }
        tmem_st16(tS + c * 16, pk);
      };
      auto row_max = [&]() {
        float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(s[c][i]));
        return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      };
      // ---- conditional rescale: the reference max only moves when the row max grows by more
      //      than 2^kRescaleThreshold; below that P <= 2^threshold, safe in fp32 sums and 16-bit P
      const float m_tile = row_max();
      if (j == 0) m_run = m_tile;
      FA_TRACE((warp & 3) == 0 && lane == 0, j, t * 5 + 2);
      float acc_scale = 1.f;
      const float grow_log2 = (m_tile - m_run) * p.scale_log2;  // 0 at j == 0
      const bool moved = grow_log2 > kRescaleThreshold;
      if (__any_sync(0xffffffffu, moved)) {
        if (moved) {
          acc_scale = ex2(-grow_log2);
          m_run = m_tile;
        }
        // PV_t(j-1) completed before S_t(j) (in-order tensor pipe), PV_t(j) waits for
        // p_full below: O_t is quiescent here.
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + c * 32, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * acc_scale);
          tmem_st32(tO + c * 32, o);
        }
      }
      // P is handed to the MMA warp in kParts pieces: the first k-steps of O += P V run while the
      // later exponentials are still being computed
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        exp_chunk(c, m_run);
        if ((c + 1) % (4 / kParts) == 0) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive_warp(&p_full[4 * t + (c + 1) / (4 / kParts) - 1]);
          FA_TRACE((warp & 3) == 0 && lane == 0, j, t * 5 + 3 + (c == 3));
        }
      }
      const uint64_t st2 = add_f32x2(sum2[0], sum2[1]);
      l_run = l_run * acc_scale + (lo_f32(st2) + hi_f32(st2));
    }

#ifdef FA_FWD_TRACE
    if (p.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.prof[262] = clock64();
#endif
    // Normalise the accumulator in TMEM and store it.  kMergeNone / kMergeLast write the 16-bit O and L;
    // kMergeFirst / kMergeMiddle write the running fp32 pair; Middle / Last first fold the running pair in.
    auto store_rows = [&](float m_fin, float l_fin) {
      const int mode = grow < p.half_rows ? p.merge_lo : p.merge_hi;
      const int64_t head_off = (int64_t)b * p.batch_stride + (int64_t)h * p.head_stride;
      const int64_t li = head_off / D + grow;
      const bool in_range = grow < p.Nq;
      float lse = m_fin * p.scale + lg2(l_fin) * kLn2;
      float w_part = 1.f / l_fin, w_acc = 0.f;
      if (mode >= kMergeMiddle && in_range) {
        const float la = p.L_acc[li];
        const float mx = fmaxf(la, lse);
        const float ea = ex2((la - mx) * kLog2e), ep = ex2((lse - mx) * kLog2e);
        const float l_new = mx + lg2(ea + ep) * kLn2;
        w_acc = ex2((la - l_new) * kLog2e);
        w_part *= ex2((lse - l_new) * kLog2e);
        lse = l_new;
      }
      const bool to_acc = mode == kMergeFirst || mode == kMergeMiddle;
      uint16_t *orow = reinterpret_cast<uint16_t *>(p.O) + head_off + (int64_t)grow * D;
      float *arow = p.O_acc + head_off + (int64_t)grow * D;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(tO + c * 32, o);
        tmem_wait_ld();
        if (!in_range) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(o[i]) * w_part;
        if (mode >= kMergeMiddle) {
          const float4 *a4 = reinterpret_cast<const float4 *>(arow + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a = a4[i];
            v[4 * i] = fmaf(a.x, w_acc, v[4 * i]); v[4 * i + 1] = fmaf(a.y, w_acc, v[4 * i + 1]);
            v[4 * i + 2] = fmaf(a.z, w_acc, v[4 * i + 2]); v[4 * i + 3] = fmaf(a.w, w_acc, v[4 * i + 3]);
          }
        }
        if (to_acc) {
          float4 *d4 = reinterpret_cast<float4 *>(arow + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
          uint4 *dst = reinterpret_cast<uint4 *>(orow + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(pack2<IS_BF16>(v[8 * i], v[8 * i + 1]), pack2<IS_BF16>(v[8 * i + 2], v[8 * i + 3]),
                                pack2<IS_BF16>(v[8 * i + 4], v[8 * i + 5]), pack2<IS_BF16>(v[8 * i + 6], v[8 * i + 7]));
        }
      }
      if (in_range) {
        if (to_acc) p.L_acc[li] = lse;
        else if (p.L != nullptr) p.L[li] = lse;
      }
    };
    if (split == 1) {
      if (nt > 0) {
        // ------------------------------ epilogue -------------------------------
        mbar_wait(&o_full[t], 0);
        tc_fence_after();
        store_rows(m_run, l_run);
      }
    } else {
      // ------------------------- epilogue of a key split ----------------------
      // Binary tree over the cluster ranks: in round `step` rank r + step sends its partial
      // (m, l, unnormalised O) to rank r (r a multiple of 2 * step) through distributed shared
      // memory; the receiver merges by the online-softmax rule and keeps the result in its own
      // TMEM accumulator; after the last round rank 0 normalises and stores.
      if (nt > 0) {
        mbar_wait(&o_full[t], 0);
        tc_fence_after();
      }
      float *xO = reinterpret_cast<float *>(sKV);  // [tile][column][row]: conflict-free both ways
      float *xM = reinterpret_cast<float *>(sQ);   // [tile][row]
      float *xL = xM + 2 * kBM;
      float m_cur = nt > 0 ? m_run : -CUDART_INF_F, l_cur = nt > 0 ? l_run : 0.f;
      bool o_valid = nt > 0;  // my TMEM accumulator holds data
      for (int step = 1; step < split; step *= 2) {
        const bool sender = (crank & (2 * step - 1)) == step, receiver = (crank & (2 * step - 1)) == 0;
        cluster_arrive();  // the receivers' K/V stages and Q tiles are dead (tile loop over, or the
        cluster_wait();    // previous round has been read) and can take a partial result
        if (sender) {
          const uint32_t dst = (uint32_t)(crank - step);
          const uint32_t rO = mapa_shared(smem_u32(xO), dst), rM = mapa_shared(smem_u32(xM), dst),
                         rL = mapa_shared(smem_u32(xL), dst);
          st_cluster_f32(rM + (t * kBM + row_in_tile) * 4, m_cur);
          st_cluster_f32(rL + (t * kBM + row_in_tile) * 4, l_cur);
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            if (o_valid) {
              tmem_ld32(tO + c * 32, o);
              tmem_wait_ld();
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i)
              st_cluster_f32(rO + ((t * D + c * 32 + i) * kBM + row_in_tile) * 4, __uint_as_float(o[i]));
          }
        }
        cluster_arrive();  // the senders' stores are visible to the receivers
        cluster_wait();
        if (receiver) {
          const float m_b = xM[t * kBM + row_in_tile], l_b = xL[t * kBM + row_in_tile];
          const float m = fmaxf(m_cur, m_b);
          const float w_a = m_cur == -CUDART_INF_F ? 0.f : ex2((m_cur - m) * p.scale_log2);
          const float w_b = m_b == -CUDART_INF_F ? 0.f : ex2((m_b - m) * p.scale_log2);
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            if (o_valid) {
              tmem_ld32(tO + c * 32, o);
              tmem_wait_ld();
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i)
              o[i] = __float_as_uint(__uint_as_float(o[i]) * w_a + xO[(t * D + c * 32 + i) * kBM + row_in_tile] * w_b);
            tmem_st32(tO + c * 32, o);
          }
          tmem_wait_st();
          l_cur = l_cur * w_a + l_b * w_b;
          m_cur = m;
          o_valid = true;
        }
      }
      if (crank == 0 && n_t[t] > 0) store_rows(m_cur, l_cur);
    }
  } else {
    setmaxnreg_dec<64>();
    if (warp == kLoadWarp) {
      // ============================== TMA producer ==============================
      if (elect_one()) {
        prefetch_tensormap(&tmQ);
        prefetch_tensormap(&tmK);
        prefetch_tensormap(&tmV);
#pragma unroll
        for (int t = 0; t < 2; ++t)
          if (c_t[t] > 0) {
            mbar_arrive_expect_tx(&q_full[t], Cfg::kTileBytes);
#pragma unroll
            for (int c = 0; c < Cfg::kChunks; ++c)
              tma_load_4d(sQ + t * Cfg::kTileBytes + c * Cfg::kChunkBytes, &tmQ, &q_full[t], c * 64,
                          q_row0 + t * kBM, h, b);
          }
        for (int item = 0; item < 2 * nmax; ++item) {
          const int stage = item % Cfg::kStages;
          const int round = item / Cfg::kStages;
          mbar_wait(&kv_empty[stage], (round & 1) ^ 1);
          FA_TRACE(true, (item >> 1), 40 + (item & 1));
          mbar_arrive_expect_tx(&kv_full[stage], Cfg::kTileBytes);
          const CUtensorMap *map = (item & 1) ? &tmV : &tmK;
#pragma unroll
          for (int c = 0; c < Cfg::kChunks; ++c)
            tma_load_4d(sKV + stage * Cfg::kTileBytes + c * Cfg::kChunkBytes, map, &kv_full[stage],
                        c * 64, (j_begin + (item >> 1)) * kBN, h, b);
        }
      }
      __syncwarp();
    } else if (warp == kMmaWarp) {
      // =============================== MMA issuer ===============================
      if (nmax > 0 && elect_one()) {  // (a rank of a key split can be left without tiles on short causal rows)
        constexpr uint32_t idesc_qk = make_idesc(kBM, kBN, IS_BF16, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc(kBM, D, IS_BF16, 0, 1);
        const uint32_t sQ_addr = smem_u32(sQ);
        const uint32_t sKV_addr = smem_u32(sKV);
        auto wait_full = [&](int item) -> uint32_t {
          const int stage = item % Cfg::kStages;
          mbar_wait(&kv_full[stage], (item / Cfg::kStages) & 1);
          tc_fence_after();
          return (uint32_t)stage;
        };
        auto issue_qk = [&](int t, uint32_t kstage) {
          const uint32_t qa = sQ_addr + t * Cfg::kTileBytes;
          const uint32_t ka = sKV_addr + kstage * Cfg::kTileBytes;
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            const uint32_t off = (kk >> 2) * Cfg::kChunkBytes + (kk & 3) * 32;
            mma_ss(tmem_base + t * kBN, make_sdesc_sw128(qa + off, 16, 1024),
                   make_sdesc_sw128(ka + off, 16, 1024), idesc_qk, kk > 0);
          }
        };
        constexpr int kParts = D == 64 ? FA_FWD_PARTS64 : FA_FWD_PARTS128;
        auto issue_pv = [&](int t, int part, uint32_t vstage, bool accumulate) {
          const uint32_t va = sKV_addr + vstage * Cfg::kTileBytes;
#pragma unroll
          for (int kk = part * (8 / kParts); kk < (part + 1) * (8 / kParts); ++kk)
            mma_ts(tmem_base + 256 + t * D, tmem_base + t * kBN + kk * 8,
                   make_sdesc_sw128(va + kk * 2048, Cfg::kChunkBytes, 1024), idesc_pv,
                   (accumulate || kk > 0) ? 1u : 0u);
        };

        const uint32_t ks0 = wait_full(0);
#pragma unroll
        for (int t = 0; t < 2; ++t)
          if (c_t[t] > 0) {
            mbar_wait(&q_full[t], 0);
            tc_fence_after();
            issue_qk(t, ks0);
            tc_commit(&s_full[t]);
          }
        tc_commit(&kv_empty[ks0]);
        for (int j = 0; j < nmax; ++j) {
          FA_TRACE(true, j, 18);
          const uint32_t vs = wait_full(2 * j + 1);
          FA_TRACE(true, j, 19);
          const bool more = j + 1 < nmax;
          const uint32_t ks = more ? wait_full(2 * j + 2) : 0u;
          FA_TRACE(true, j, 20);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (j < c_t[t]) {
#pragma unroll
              for (int part = 0; part < kParts; ++part) {
                mbar_wait(&p_full[4 * t + part], j & 1);
                tc_fence_after();
                FA_TRACE(part == 0 || part == kParts - 1, j, 21 + t * 5 + (part > 0));
                issue_pv(t, part, vs, j > 0);
              }
              FA_TRACE(true, j, 23 + t * 5);
              if (j == c_t[t] - 1) tc_commit(&o_full[t]);
            }
            if (j + 1 < c_t[t]) {
              issue_qk(t, ks);
              tc_commit(&s_full[t]);
              FA_TRACE(true, j, 25 + t * 5);
            }
          }
          tc_commit(&kv_empty[vs]);
          if (more) tc_commit(&kv_empty[ks]);
        }
      }
      __syncwarp();
    }
  }

  if (split > 1 && warp >= 8) {  // the cluster barriers of the split epilogue are for every thread
    for (int step = 1; step < split; step *= 2) {
      cluster_arrive();
      cluster_wait();
      cluster_arrive();
      cluster_wait();
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef FA_FWD_TRACE
  if (p.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.prof[263] = clock64();
#endif
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
#ifdef FA_FWD_TRACE
  if (p.prof != nullptr && threadIdx.x == 0 && blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.prof[258] = clock64();
    p.prof[259] = (long long)ns;
  }
#endif
}

// largest cluster of these one-CTA-per-SM blocks the current device can co-schedule (asked once per device)
template <int D, int IS_BF16>
int max_cluster_size() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  int best = cache[dev].load(std::memory_order_relaxed);
  if (best != 0) return best;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = FwdCfg<D>::kSmemBytes;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  best = 1;
  for (int c = 8; c > 1; c /= 2) {
    int n_clusters = 0;
    cfg.gridDim = dim3((unsigned)c, 1, 1);
    attr[0].val.clusterDim.x = (unsigned)c;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, fwd_tc_kernel<D, IS_BF16>, &cfg) == cudaSuccess && n_clusters > 0) {
      best = c;
      break;
    }
    (void)cudaGetLastError();
  }
  cache[dev].store(best, std::memory_order_relaxed);
  return best;
}

// Upper bound on the key-split cluster size.  4 by default: clusters of 8 work but measured slower
// (single head N=4096: 56 us unsplit, 37 / 28 / 42 us at 2 / 4 / 8: a third merge round, and eight
// whole-SM CTAs have to be co-scheduled in one GPC).  fa_debug_set_fwd_split_max changes it (tests).
std::atomic<int> g_split_cap{4};

// per-device one-time set-up of one instantiation (also what fa_preload_kernels runs ahead of time)
template <int D, int IS_BF16>
int configure_fwd_tc() {
  static DeviceOnce configured;  // the attribute is per device
  return configured.run([] {
    FA_CUDA_CHECK(cudaFuncSetAttribute(fwd_tc_kernel<D, IS_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       FwdCfg<D>::kSmemBytes));
    cudaFuncAttributes attr;  // forces the (lazily loaded) kernel into the context
    FA_CUDA_CHECK(cudaFuncGetAttributes(&attr, fwd_tc_kernel<D, IS_BF16>));
    return (int)FA_OK;
  });
}

template <int D, int IS_BF16>
int launch_fwd_tc_impl(const CUtensorMap &tmQ, const CUtensorMap &tmK, const CUtensorMap &tmV,
                       const FwdParams &p, int B, cudaStream_t stream) {
  using Cfg = FwdCfg<D>;
  int rc = configure_fwd_tc<D, IS_BF16>();
  if (rc != FA_OK) return rc;
  FwdParams q = p;
  q.n_blocks = (p.Nq + 2 * kBM - 1) / (2 * kBM);
  q.n_heads = B * p.H;
  q.group = dispatch_group(p.causal != 0, (int64_t)2 * p.Nk * D * 2, q.n_heads);
  if (q.n_blocks > 65535) q.group = 1;  // grid.y limit of the grouped form
  // fewer row blocks than half the SMs: several CTAs (one cluster) per row block, splitting the keys:
  // largest power of two that still fits one wave and leaves every rank at least four key tiles of
  // the longest row.  A ring-attention partial (merge mode set) is never split: its epilogue already
  // merges with the running result.
  const int n_sm = device_sm_count();
  q.split = 1;
  const int split_cap = (p.merge_lo | p.merge_hi) ? 1 : g_split_cap.load(std::memory_order_relaxed);
  while (q.split < split_cap && 2 * q.split * q.n_blocks * q.n_heads <= n_sm && (p.Nk + kBN - 1) / kBN >= 4 * q.split) q.split *= 2;
  if (q.split > 1) q.split = std::min(q.split, max_cluster_size<D, IS_BF16>());
  if (q.split > 1) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)q.split;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    q.group = 1;
    cfg.gridDim = dim3((unsigned)(q.n_blocks * q.split), (unsigned)p.H, (unsigned)B);
    FA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, fwd_tc_kernel<D, IS_BF16>, tmQ, tmK, tmV, q));
  } else {
    fwd_tc_kernel<D, IS_BF16><<<dispatch_grid(q.group, q.n_blocks, p.H, B), kThreads, Cfg::kSmemBytes, stream>>>(tmQ, tmK, tmV, q);
  }
  FA_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return FA_OK;
}

}  // namespace

int preload_fwd_tc() {
  int rc;
  if ((rc = configure_fwd_tc<64, 0>()) || (rc = configure_fwd_tc<64, 1>()) || (rc = configure_fwd_tc<128, 0>()) ||
      (rc = configure_fwd_tc<128, 1>()))
    return rc;
  return FA_OK;
}

void set_fwd_split_max(int cap) { g_split_cap.store(cap < 1 ? 1 : (cap > 8 ? 8 : cap), std::memory_order_relaxed); }

int launch_fwd_tc_rect(const void *Q, const void *K, const void *V, void *O, float *L, int Nq, int Nk,
                       int D, float scale, int64_t q_batch_stride, int64_t q_head_stride,
                       int64_t kv_batch_stride, int64_t kv_head_stride, int is_causal, int B, int H,
                       int dtype, cudaStream_t stream, const FwdMerge *merge) {
  FA_REQUIRE(Q && K && V && O, "null tensor pointer");
  FA_REQUIRE(Nq >= 1 && Nk >= 1, "N must be >= 1 (got %d x %d)", Nq, Nk);
  FA_REQUIRE(!is_causal || Nq == Nk, "causal attention needs Nq == Nk");
  FA_REQUIRE(D == 64 || D == 128, "D must be 64 or 128 (got %d)", D);
  FA_REQUIRE(B >= 1 && H >= 1 && H <= 65535 && B <= 65535, "bad B/H (%d, %d)", B, H);
  FA_REQUIRE(dtype == FA_DTYPE_FP16 || dtype == FA_DTYPE_BF16, "dtype must be FA_DTYPE_FP16 or FA_DTYPE_BF16");
  FA_REQUIRE(scale > 0.f, "scale must be positive");
  FA_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(O), "Q/K/V/O must be 16-byte aligned");
  FA_REQUIRE(q_batch_stride % 8 == 0 && q_head_stride % 8 == 0 && kv_batch_stride % 8 == 0 && kv_head_stride % 8 == 0,
             "strides must be multiples of 8 elements");
  FA_REQUIRE((H == 1 || (q_head_stride >= (int64_t)Nq * D && kv_head_stride >= (int64_t)Nk * D)) &&
                 (B == 1 || (q_batch_stride >= (int64_t)Nq * D && kv_batch_stride >= (int64_t)Nk * D)),
             "heads overlap: stride smaller than N*D");
  FA_REQUIRE(L == nullptr || (q_head_stride % D == 0 && q_batch_stride % D == 0),
             "L_out needs strides that are multiples of D (L index = offset / D, kernels.metal:623)");
  const CUtensorMap *tmQ, *tmK, *tmV;
  int rc;
  if ((rc = tensor_map_bhnd(&tmQ, Q, dtype, Nq, D, H, B, q_head_stride, q_batch_stride, kBM)) != FA_OK) return rc;
  if ((rc = tensor_map_bhnd(&tmK, K, dtype, Nk, D, H, B, kv_head_stride, kv_batch_stride, kBN)) != FA_OK) return rc;
  if ((rc = tensor_map_bhnd(&tmV, V, dtype, Nk, D, H, B, kv_head_stride, kv_batch_stride, kBN)) != FA_OK) return rc;
  FwdParams p = {};
#ifdef FA_FWD_TRACE
  p.prof = g_trace_buffer;
#endif
  p.O = O;
  p.L = L;
  p.Nq = Nq;
  p.Nk = Nk;
  p.H = H;
  p.scale = scale;
  p.scale_log2 = scale * kLog2e;
  p.batch_stride = q_batch_stride;
  p.head_stride = q_head_stride;
  p.causal = is_causal ? 1 : 0;
  if (merge != nullptr && (merge->lo != kMergeNone || merge->hi != kMergeNone)) {
    FA_REQUIRE(merge->O_acc && merge->L_acc && aligned16(merge->O_acc), "merge needs 16-byte aligned O_acc and L_acc");
    FA_REQUIRE(q_head_stride % D == 0 && q_batch_stride % D == 0, "merge needs strides that are multiples of D");
    p.O_acc = merge->O_acc;
    p.L_acc = merge->L_acc;
    p.merge_lo = merge->lo;
    p.merge_hi = merge->hi;
    p.half_rows = merge->half_rows;
  }
  if (D == 64)
    return dtype == FA_DTYPE_BF16 ? launch_fwd_tc_impl<64, 1>(*tmQ, *tmK, *tmV, p, B, stream)
                                  : launch_fwd_tc_impl<64, 0>(*tmQ, *tmK, *tmV, p, B, stream);
  return dtype == FA_DTYPE_BF16 ? launch_fwd_tc_impl<128, 1>(*tmQ, *tmK, *tmV, p, B, stream)
                                : launch_fwd_tc_impl<128, 0>(*tmQ, *tmK, *tmV, p, B, stream);
}

int launch_fwd_tc(const void *Q, const void *K, const void *V, void *O, float *L, int N, int D,
                  float scale, int64_t batch_stride, int64_t head_stride, int is_causal, int B, int H,
                  int dtype, cudaStream_t stream) {
  return launch_fwd_tc_rect(Q, K, V, O, L, N, N, D, scale, batch_stride, head_stride, batch_stride, head_stride,
                            is_causal, B, H, dtype, stream, nullptr);
}

}  // namespace fa
