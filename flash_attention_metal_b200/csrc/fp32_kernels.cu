// fp32 attention variants for sm_100a: naive, tiled (V1) and vectorised (V2).
//
// These mirror the three fp32 kernels of the reference (kernels.metal:12-64,
// 72-171, 462-596) in *role*, not in code: "naive" stays a one-thread-per-row
// two-pass kernel because it is the denominator of the CSV's speed-up columns;
// V1 is the straightforward shared-memory tiling; V2 is the performance path --
// 128-bit loads, cp.async double buffering, a 4x4 register tile per thread and
// half-warp shuffles for the online-softmax statistics.  None of them touches
// tensor cores (fp32 inputs); their ceiling is the FFMA pipe.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fa_internal.h"
#include "sm100_ptx.cuh"

namespace fa {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ int64_t head_offset(int64_t batch_stride, int64_t head_stride) {
  return (int64_t)blockIdx.z * batch_stride + (int64_t)blockIdx.y * head_stride;
}

// ---------------------------------------------------------------------------
// naive: one thread per query row, two passes over all keys, everything read
// from global memory (role of kernels.metal:12-64).
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) naive_attention_kernel(
    const float *__restrict__ Q, const float *__restrict__ K, const float *__restrict__ V,
    float *__restrict__ O, int N, float scale, int is_causal, int64_t batch_stride,
    int64_t head_stride) {
  const int64_t off = head_offset(batch_stride, head_stride);
  Q += off; K += off; V += off; O += off;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int nk = is_causal ? i + 1 : N;
  const float4 *q4 = reinterpret_cast<const float4 *>(Q + (int64_t)i * D);

  float max_score = -CUDART_INF_F;
  for (int j = 0; j < nk; ++j) {
    const float4 *k4 = reinterpret_cast<const float4 *>(K + (int64_t)j * D);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      float4 a = q4[c], b = __ldg(k4 + c);
      s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    max_score = fmaxf(max_score, s * scale);
  }
  float acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.f;
  float sum_exp = 0.f;
  for (int j = 0; j < nk; ++j) {
    const float4 *k4 = reinterpret_cast<const float4 *>(K + (int64_t)j * D);
    const float4 *v4 = reinterpret_cast<const float4 *>(V + (int64_t)j * D);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      float4 a = q4[c], b = __ldg(k4 + c);
      s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    const float p = expf(s * scale - max_score);
    sum_exp += p;
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      float4 b = __ldg(v4 + c);
      acc[4 * c + 0] += p * b.x;
      acc[4 * c + 1] += p * b.y;
      acc[4 * c + 2] += p * b.z;
      acc[4 * c + 3] += p * b.w;
    }
  }
  const float inv = 1.f / sum_exp;
  float4 *o4 = reinterpret_cast<float4 *>(O + (int64_t)i * D);
#pragma unroll
  for (int c = 0; c < D / 4; ++c)
    o4[c] = make_float4(acc[4 * c] * inv, acc[4 * c + 1] * inv, acc[4 * c + 2] * inv,
                        acc[4 * c + 3] * inv);
}

// ---------------------------------------------------------------------------
// V1: shared-memory tiling, one thread per query row, Br = 64 rows per block,
// Bc = 32 keys per tile, online softmax rescaled once per tile (role of
// kernels.metal:72-171; the reference rescales once per *key*).
// ---------------------------------------------------------------------------
constexpr int V1_BR = 64;
constexpr int V1_BC = 32;

template <int D>
__global__ void __launch_bounds__(V1_BR) flash_attention_v1_kernel(
    const float *__restrict__ Q, const float *__restrict__ K, const float *__restrict__ V,
    float *__restrict__ O, int N, float scale, int is_causal, int64_t batch_stride,
    int64_t head_stride) {
  __shared__ float Ks[V1_BC][D];
  __shared__ float Vs[V1_BC][D];
  const int64_t off = head_offset(batch_stride, head_stride);
  Q += off; K += off; V += off; O += off;
  const int tx = threadIdx.x;
  const int row0 = blockIdx.x * V1_BR;
  const int i = row0 + tx;
  const bool live = i < N;

  float q[D], o[D];
#pragma unroll
  for (int c = 0; c < D / 4; ++c) {
    float4 a = live ? reinterpret_cast<const float4 *>(Q + (int64_t)i * D)[c]
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    q[4 * c] = a.x * scale; q[4 * c + 1] = a.y * scale; q[4 * c + 2] = a.z * scale; q[4 * c + 3] = a.w * scale;
  }
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = 0.f;
  float m = -CUDART_INF_F, l = 0.f;

  const int last_row = min(row0 + V1_BR, N) - 1;
  const int kend = is_causal ? last_row + 1 : N;
  for (int j0 = 0; j0 < kend; j0 += V1_BC) {
    __syncthreads();
    // cooperative, coalesced 128-bit tile load; keys past N are zero-filled and masked below
    for (int idx = tx; idx < V1_BC * D / 4; idx += V1_BR) {
      const int r = idx / (D / 4), c = idx % (D / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j0 + r < N) {
        kv = __ldg(reinterpret_cast<const float4 *>(K + (int64_t)(j0 + r) * D) + c);
        vv = __ldg(reinterpret_cast<const float4 *>(V + (int64_t)(j0 + r) * D) + c);
      }
      reinterpret_cast<float4 *>(&Ks[r][0])[c] = kv;
      reinterpret_cast<float4 *>(&Vs[r][0])[c] = vv;
    }
    __syncthreads();
    float s[V1_BC];
    float tile_max = -CUDART_INF_F;
#pragma unroll
    for (int r = 0; r < V1_BC; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) acc += q[d] * Ks[r][d];
      const int j = j0 + r;
      const bool masked = (j >= N) || (is_causal && j > i);
      s[r] = masked ? -CUDART_INF_F : acc;
      tile_max = fmaxf(tile_max, s[r]);
    }
    const float m_new = fmaxf(m, tile_max);
    if (m_new == -CUDART_INF_F) continue;  // nothing visible yet for this row (uniform sync above)
    const float corr = expf(m - m_new);
    l *= corr;
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] *= corr;
#pragma unroll
    for (int r = 0; r < V1_BC; ++r) {
      const float p = expf(s[r] - m_new);
      l += p;
#pragma unroll
      for (int d = 0; d < D; ++d) o[d] += p * Vs[r][d];
    }
    m = m_new;
  }
  if (live) {
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < D / 4; ++c)
      reinterpret_cast<float4 *>(O + (int64_t)i * D)[c] =
          make_float4(o[4 * c] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv);
  }
}

// ---------------------------------------------------------------------------
// V2: the fp32 performance path (role of kernels.metal:462-596).
//   block  = 256 threads = 16 (ty) x 16 (tx), BM = 64 query rows, BN = 64 keys
//   S tile : thread (ty, tx) owns rows ty+16i, keys tx+16j (4x4 register tile),
//            float4 dot products from padded shared memory (conflict-free)
//   softmax: row statistics reduced across the 16 lanes of a half warp by
//            shuffles; exp2 with log2(e) folded into the scale
//   PV     : P goes through shared memory (only the owning half warp reads its
//            rows back -> __syncwarp suffices); thread owns rows ty+16i and
//            columns tx*4 (+64 at D=128) of O
//   K/V    : cp.async 16-byte copies, two stages
// ---------------------------------------------------------------------------
constexpr int V2_BM = 64;
constexpr int V2_BN = 64;
constexpr int V2_THREADS = 256;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int bytes = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int D>
struct V2Smem {
  static constexpr int LD = D + 4;  // row pitch in floats: 16B aligned, conflict-free float4 reads
  float q[V2_BM][LD];
  float k[2][V2_BN][LD];
  float v[2][V2_BN][LD];
  float p[V2_BM][V2_BN + 4];
};

template <int D>
__global__ void __launch_bounds__(V2_THREADS) flash_attention_v2_kernel(
    const float *__restrict__ Q, const float *__restrict__ K, const float *__restrict__ V,
    float *__restrict__ O, int N, float scale, int is_causal, int64_t batch_stride,
    int64_t head_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V2Smem<D> &sm = *reinterpret_cast<V2Smem<D> *>(smem_raw);
  constexpr int C4 = D / 4;     // float4 per row
  constexpr int OC = D / 64;    // float4 output columns per thread

  const int64_t off = head_offset(batch_stride, head_stride);
  Q += off; K += off; V += off; O += off;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // heaviest (latest) row blocks first when causal
  const int bx = is_causal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int row0 = bx * V2_BM;
  const float scale_log2 = scale * kLog2e;

  auto load_kv = [&](int stage, int j0) {
    for (int idx = tid; idx < V2_BN * C4; idx += V2_THREADS) {
      const int r = idx / C4, c = idx % C4;
      const bool ok = j0 + r < N;
      const int64_t g = (int64_t)(ok ? j0 + r : 0) * D + 4 * c;
      cp_async16(&sm.k[stage][r][4 * c], K + g, ok);
      cp_async16(&sm.v[stage][r][4 * c], V + g, ok);
    }
  };

  const int last_row = min(row0 + V2_BM, N) - 1;
  const int kend = is_causal ? last_row + 1 : N;
  const int ntiles = (kend + V2_BN - 1) / V2_BN;

  for (int idx = tid; idx < V2_BM * C4; idx += V2_THREADS) {
    const int r = idx / C4, c = idx % C4;
    const bool ok = row0 + r < N;
    cp_async16(&sm.q[r][4 * c], Q + (int64_t)(ok ? row0 + r : 0) * D + 4 * c, ok);
  }
  load_kv(0, 0);
  cp_async_commit();

  uint64_t o2[4][OC * 2];  // output accumulators as fp32x2 pairs: columns (4 oc, 4 oc + 1) and (4 oc + 2, 4 oc + 3)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < OC * 2; ++c) o2[i][c] = 0ull;
  float m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -CUDART_INF_F; l[i] = 0.f; }

  for (int t = 0; t < ntiles; ++t) {
    const int st = t & 1, j0 = t * V2_BN;
    if (t + 1 < ntiles) load_kv(st ^ 1, j0 + V2_BN);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    // ---- S = Q K^T on a 4x4 register tile --------------------------------
    // packed fp32x2 FMAs (FFMA2: one issue slot for two lanes): each accumulator is a pair of partial
    // sums over the even / odd pairs of the head dimension, added at the end.  Halves the issue slots of
    // the FMA-bound loops so the FMA pipe, not the scheduler, is the limit.
    float s[4][4];
    {
      uint64_t s2[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s2[i][j] = 0ull;
      const uint32_t q_base = (uint32_t)__cvta_generic_to_shared(&sm.q[ty][0]);
      const uint32_t k_base = (uint32_t)__cvta_generic_to_shared(&sm.k[st][tx][0]);
      constexpr uint32_t kRow16 = 16u * V2Smem<D>::LD * 4u;  // byte distance of rows 16 apart
#pragma unroll 4
      for (int c = 0; c < C4; ++c) {
        uint64_t a[4][2], b[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) ptx::lds_v2b64(q_base + i * kRow16 + c * 16, a[i][0], a[i][1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) ptx::lds_v2b64(k_base + j * kRow16 + c * 16, b[j][0], b[j][1]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s2[i][j] = ptx::fma_f32x2(a[i][0], b[j][0], s2[i][j]);
            s2[i][j] = ptx::fma_f32x2(a[i][1], b[j][1], s2[i][j]);
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = ptx::lo_f32(s2[i][j]) + ptx::hi_f32(s2[i][j]);
    }
    // ---- mask + online softmax -------------------------------------------
    const bool need_mask = (j0 + V2_BN > N) || (is_causal && j0 + V2_BN - 1 > row0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = row0 + ty + 16 * i;
      float tmax = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gj = j0 + tx + 16 * j;
        if (need_mask && (gj >= N || (is_causal && gj > gi))) s[i][j] = -CUDART_INF_F;
        tmax = fmaxf(tmax, s[i][j]);
      }
#pragma unroll
      for (int w = 8; w >= 1; w >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, w));
      const float m_new = fmaxf(m[i], tmax);
      // rows that see no key at all in this tile and before keep m = -inf: guard the subtraction
      const float m_ref = (m_new == -CUDART_INF_F) ? 0.f : m_new;
      const float corr = exp2f((m[i] - m_ref) * scale_log2);
      m[i] = m_new;
      l[i] *= corr;
      const uint64_t corr2 = ptx::pack_f32x2(corr, corr);
#pragma unroll
      for (int c = 0; c < OC * 2; ++c) o2[i][c] = ptx::mul_f32x2(o2[i][c], corr2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = exp2f((s[i][j] - m_ref) * scale_log2);
        l[i] += p;
        sm.p[ty + 16 * i][tx + 16 * j] = p;
      }
    }
    __syncwarp();
    // ---- O += P V ----------------------------------------------------------
    {
      const uint32_t p_base = (uint32_t)__cvta_generic_to_shared(&sm.p[ty][0]);
      const uint32_t v_base = (uint32_t)__cvta_generic_to_shared(&sm.v[st][0][4 * tx]);
      constexpr uint32_t kPRow16 = 16u * (V2_BN + 4) * 4u, kVRow = V2Smem<D>::LD * 4u;
#pragma unroll 2
      for (int k4 = 0; k4 < V2_BN / 4; ++k4) {
        float4 p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint64_t lo, hi;
          ptx::lds_v2b64(p_base + i * kPRow16 + k4 * 16, lo, hi);
          p[i] = make_float4(ptx::lo_f32(lo), ptx::hi_f32(lo), ptx::lo_f32(hi), ptx::hi_f32(hi));
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint64_t pp[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float pv = kk == 0 ? p[i].x : kk == 1 ? p[i].y : kk == 2 ? p[i].z : p[i].w;
            pp[i] = ptx::pack_f32x2(pv, pv);
          }
#pragma unroll
          for (int oc = 0; oc < OC; ++oc) {
            uint64_t v01, v23;
            ptx::lds_v2b64(v_base + (4 * k4 + kk) * kVRow + oc * 256, v01, v23);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              o2[i][2 * oc] = ptx::fma_f32x2(pp[i], v01, o2[i][2 * oc]);
              o2[i][2 * oc + 1] = ptx::fma_f32x2(pp[i], v23, o2[i][2 * oc + 1]);
            }
          }
        }
      }
    }
    __syncthreads();  // stage st is refilled by the next iteration's prefetch
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lt = l[i];
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) lt += __shfl_xor_sync(0xffffffffu, lt, w);
    const int gi = row0 + ty + 16 * i;
    if (gi < N) {
      const float inv = 1.f / lt;
#pragma unroll
      for (int oc = 0; oc < OC; ++oc)
        *reinterpret_cast<float4 *>(O + (int64_t)gi * D + 64 * oc + 4 * tx) =
            make_float4(ptx::lo_f32(o2[i][2 * oc]) * inv, ptx::hi_f32(o2[i][2 * oc]) * inv,
                        ptx::lo_f32(o2[i][2 * oc + 1]) * inv, ptx::hi_f32(o2[i][2 * oc + 1]) * inv);
    }
  }
}

template <int D>
int launch_fp32_d(int variant, const float *Q, const float *K, const float *V, float *O, int N,
                  float scale, int64_t bs, int64_t hs, int causal, int B, int H, cudaStream_t st) {
  if (variant == 0) {
    dim3 grid((N + 127) / 128, H, B);
    naive_attention_kernel<D><<<grid, 128, 0, st>>>(Q, K, V, O, N, scale, causal, bs, hs);
  } else if (variant == 1) {
    dim3 grid((N + V1_BR - 1) / V1_BR, H, B);
    flash_attention_v1_kernel<D><<<grid, V1_BR, 0, st>>>(Q, K, V, O, N, scale, causal, bs, hs);
  } else {
    const int smem = (int)sizeof(V2Smem<D>);
    static DeviceOnce configured;  // the attribute is per device
    const int rc = configured.run([smem] {
      FA_CUDA_CHECK(cudaFuncSetAttribute(flash_attention_v2_kernel<D>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      return (int)FA_OK;
    });
    if (rc != FA_OK) return rc;
    dim3 grid((N + V2_BM - 1) / V2_BM, H, B);
    flash_attention_v2_kernel<D><<<grid, V2_THREADS, smem, st>>>(Q, K, V, O, N, scale, causal, bs, hs);
  }
  FA_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return FA_OK;
}

}  // namespace

int launch_fp32(int variant, const float *Q, const float *K, const float *V, float *O, int N, int D,
                float scale, int64_t batch_stride, int64_t head_stride, int is_causal, int B, int H,
                cudaStream_t stream) {
  FA_REQUIRE(Q && K && V && O, "null tensor pointer");
  FA_REQUIRE(N >= 1, "N must be >= 1 (got %d)", N);
  FA_REQUIRE(D == 64 || D == 128, "D must be 64 or 128 (got %d)", D);
  FA_REQUIRE(B >= 1 && H >= 1 && H <= 65535 && B <= 65535, "bad B/H (%d, %d)", B, H);
  FA_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(O),
             "Q/K/V/O must be 16-byte aligned");
  FA_REQUIRE(batch_stride % 4 == 0 && head_stride % 4 == 0, "strides must be multiples of 4 elements");
  if (D == 64)
    return launch_fp32_d<64>(variant, Q, K, V, O, N, scale, batch_stride, head_stride, is_causal, B, H, stream);
  return launch_fp32_d<128>(variant, Q, K, V, O, N, scale, batch_stride, head_stride, is_causal, B, H, stream);
}

}  // namespace fa
