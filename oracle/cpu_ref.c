/*
 * oracle/cpu_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the CPU verifier that lives inline in the
 * reference's host program (main.mm).  It is the checker the CUDA kernels are
 * compared with.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product library
 * (libflash_attn_b200.so) never links, loads or calls anything in oracle/.
 *
 * Parity pin: the reference ships no golden vectors for this path (SURVEY.md
 * section 8c), and its host program is Objective-C++/Metal, so it cannot be built
 * here as a whole.  Its CPU verifier loops, however, are plain C++: the recipe
 * oracle/build_ref.sh lifts those line ranges out of /root/reference/main.mm
 * where they lie, compiles them unmodified into oracle/_ref/libref_cpu.so, and
 * tests/test_oracle.py checks every function below against that build
 * bit-for-bit (forward, causal forward, initRandom) or to 1e-6 (backward,
 * whose accumulation order differs).  The outputs of that build are frozen in
 * tests/golden/ by tests/golden/make_golden.py so the pin survives on machines
 * without /root/reference.
 *
 * Each function cites the reference lines it follows (paths are relative to
 * /root/reference).  Arithmetic is IEEE fp32, no fast-math, no FMA contraction
 * (the Makefile passes -ffp-contract=off) so results do not depend on the host
 * CPU's instruction set.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
/* Inputs: main.mm:24-30 (initRandom).                                       */
/*   std::mt19937 gen(42); uniform_real_distribution<float>(-1,1)            */
/* libstdc++ and libc++ both reduce this to  2 * (float(r) / 2^32) - 1  with  */
/* a clamp just below 1.0 for the canonical value; the generator is          */
/* re-created on every call, so every tensor filled by it is identical.      */
/* ------------------------------------------------------------------------ */
typedef struct { uint32_t mt[624]; int idx; } mt19937_t;

static void mt_seed(mt19937_t *g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i)
    g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

static uint32_t mt_next(mt19937_t *g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      uint32_t v = g->mt[(i + 397) % 624] ^ (y >> 1);
      if (y & 1u) v ^= 0x9908b0dfu;
      g->mt[i] = v;
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* main.mm:24-30 with an explicit seed (the reference always uses 42). */
void oracle_init_random_seeded(float *data, long size, uint32_t seed) {
  mt19937_t g;
  mt_seed(&g, seed);
  for (long i = 0; i < size; ++i) {
    float canon = (float)mt_next(&g) / 4294967296.0f;
    if (canon >= 1.0f) canon = nextafterf(1.0f, 0.0f);
    data[i] = canon * 2.0f + -1.0f;
  }
}

/* main.mm:24-30 exactly: seed 42, restarted on every call. */
void oracle_init_random(float *data, long size) {
  oracle_init_random_seeded(data, size, 42u);
}

/* ------------------------------------------------------------------------ */
/* fp32 <-> fp16 / bf16, round-to-nearest-even: main.mm:322-329 ((__fp16)x). */
/* The reference only has fp16; bf16 is the B200 addition (BASELINE.json).   */
/* ------------------------------------------------------------------------ */
uint16_t oracle_f32_to_f16(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t absx = x & 0x7fffffffu;
  if (absx >= 0x7f800000u) /* inf / nan */
    return (uint16_t)(sign | 0x7c00u | (absx > 0x7f800000u ? 0x200u : 0u));
  if (absx >= 0x477ff000u) /* rounds to >= 65520 -> inf */
    return (uint16_t)(sign | 0x7c00u);
  if (absx < 0x33000001u) /* < 2^-25 (or exactly 2^-25: ties to even -> 0) */
    return (uint16_t)sign;
  int exp = (int)(absx >> 23) - 127;
  uint32_t man = (absx & 0x7fffffu) | 0x800000u;
  int shift; /* how many low bits of the 24-bit significand are dropped */
  uint32_t base;
  if (exp < -14) { shift = 13 + (-14 - exp); base = 0; }
  else           { shift = 13;               base = (uint32_t)(exp + 15) << 10; }
  uint32_t kept = man >> shift;
  uint32_t rem = man & ((1u << shift) - 1u);
  uint32_t half = 1u << (shift - 1);
  if (exp >= -14) kept &= 0x3ffu; /* drop the implicit one */
  uint32_t h = base + kept;
  if (rem > half || (rem == half && (h & 1u))) ++h; /* carry may bump the exponent */
  return (uint16_t)(sign | h);
}

float oracle_f16_to_f32(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t x;
  if (exp == 0) {
    if (man == 0) x = sign;
    else {
      int e = -1;
      do { ++e; man <<= 1; } while (!(man & 0x400u));
      x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) x = sign | 0x7f800000u | (man << 13);
  else x = sign | ((exp + 112u) << 23) | (man << 13);
  float f;
  memcpy(&f, &x, 4);
  return f;
}

uint16_t oracle_f32_to_bf16(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40u);
  uint32_t lsb = (x >> 16) & 1u;
  x += 0x7fffu + lsb;
  return (uint16_t)(x >> 16);
}

float oracle_bf16_to_f32(uint16_t h) {
  uint32_t x = (uint32_t)h << 16;
  float f;
  memcpy(&f, &x, 4);
  return f;
}

/* dtype: 0 = fp16, 1 = bf16 (same codes as include/flash_attn_b200.h) */
void oracle_f32_to_half_array(const float *src, uint16_t *dst, long n, int dtype) {
  for (long i = 0; i < n; ++i)
    dst[i] = dtype ? oracle_f32_to_bf16(src[i]) : oracle_f32_to_f16(src[i]);
}

void oracle_half_to_f32_array(const uint16_t *src, float *dst, long n, int dtype) {
  for (long i = 0; i < n; ++i)
    dst[i] = dtype ? oracle_bf16_to_f32(src[i]) : oracle_f16_to_f32(src[i]);
}

/* ------------------------------------------------------------------------ */
/* Forward, "faithful": main.mm:128-159, loop order i -> d -> j -> k.        */
/* The scores are recomputed for every output column d (O(N^2 D^2)); this is */
/* BASELINE.json config 1 and is only usable up to N ~ 1024.                 */
/* ------------------------------------------------------------------------ */
void oracle_forward_faithful(const float *q, const float *k, const float *v, float *o,
                             int N, int D, float scale) {
  for (int i = 0; i < N; ++i) {
    for (int d = 0; d < D; ++d) {
      float num = 0.0f, den = 0.0f, max_score = -INFINITY;
      for (int j = 0; j < N; ++j) {
        float score = 0.0f;
        for (int kk = 0; kk < D; ++kk) score += q[(long)i * D + kk] * k[(long)j * D + kk];
        score *= scale;
        if (score > max_score) max_score = score;
      }
      for (int j = 0; j < N; ++j) {
        float score = 0.0f;
        for (int kk = 0; kk < D; ++kk) score += q[(long)i * D + kk] * k[(long)j * D + kk];
        score *= scale;
        float p = expf(score - max_score);
        num += p * v[(long)j * D + d];
        den += p;
      }
      o[(long)i * D + d] = num / den;
    }
  }
}

/* ------------------------------------------------------------------------ */
/* Forward, "hoisted": the same arithmetic per (i, d) -- identical operation */
/* order for every accumulator, hence bit-identical output -- with the score */
/* row computed once per query.  Follows main.mm:128-159 when causal == 0    */
/* and main.mm:549-578 (keys j <= i only) when causal != 0.  Also returns    */
/* L[i] = max + log(sum) as the V4 kernel defines it (kernels.metal:863).    */
/* Heads are independent (kernels.metal:622); rows run under OpenMP.         */
/* ------------------------------------------------------------------------ */
void oracle_forward(const float *q, const float *k, const float *v, float *o, float *lse,
                    int N, int D, float scale, int causal) {
#pragma omp parallel
  {
    float *scores = (float *)malloc(sizeof(float) * (size_t)N);
    float *num = (float *)malloc(sizeof(float) * (size_t)D);
#pragma omp for schedule(dynamic, 8)
    for (int i = 0; i < N; ++i) {
      int nk = causal ? i + 1 : N;
      float max_score = -INFINITY;
      for (int j = 0; j < nk; ++j) {
        float score = 0.0f;
        for (int kk = 0; kk < D; ++kk) score += q[(long)i * D + kk] * k[(long)j * D + kk];
        score *= scale;
        scores[j] = score;
        if (score > max_score) max_score = score;
      }
      float den = 0.0f;
      for (int d = 0; d < D; ++d) num[d] = 0.0f;
      for (int j = 0; j < nk; ++j) {
        float p = expf(scores[j] - max_score);
        den += p;
        for (int d = 0; d < D; ++d) num[d] += p * v[(long)j * D + d];
      }
      for (int d = 0; d < D; ++d) o[(long)i * D + d] = num[d] / den;
      if (lse) lse[i] = max_score + logf(den);
    }
    free(scores);
    free(num);
  }
}

/* Batched wrapper: contiguous [B*H, N, D] tensors, L is [B*H, N]. */
void oracle_forward_batched(const float *q, const float *k, const float *v, float *o, float *lse,
                            int heads, int N, int D, float scale, int causal) {
  for (int h = 0; h < heads; ++h) {
    long off = (long)h * N * D;
    oracle_forward(q + off, k + off, v + off, o + off, lse ? lse + (long)h * N : 0, N, D, scale,
                   causal);
  }
}

/* fp64 forward, for the oracle's own accuracy check. */
void oracle_forward_f64(const float *q, const float *k, const float *v, double *o, double *lse,
                        int N, int D, float scale, int causal) {
#pragma omp parallel
  {
    double *scores = (double *)malloc(sizeof(double) * (size_t)N);
#pragma omp for schedule(dynamic, 8)
    for (int i = 0; i < N; ++i) {
      int nk = causal ? i + 1 : N;
      double mx = -INFINITY;
      for (int j = 0; j < nk; ++j) {
        double s = 0.0;
        for (int kk = 0; kk < D; ++kk) s += (double)q[(long)i * D + kk] * (double)k[(long)j * D + kk];
        s *= (double)scale;
        scores[j] = s;
        if (s > mx) mx = s;
      }
      double den = 0.0;
      for (int j = 0; j < nk; ++j) { scores[j] = exp(scores[j] - mx); den += scores[j]; }
      for (int d = 0; d < D; ++d) {
        double acc = 0.0;
        for (int j = 0; j < nk; ++j) acc += scores[j] * (double)v[(long)j * D + d];
        o[(long)i * D + d] = acc / den;
      }
      if (lse) lse[i] = mx + log(den);
    }
    free(scores);
  }
}

/* ------------------------------------------------------------------------ */
/* Backward: main.mm:1091-1179.  Dense P = softmax(scale * Q K^T)            */
/* (:1095-1118), dV = P^T dO (:1122-1130), dP = dO V^T (:1133-1143),         */
/* dS = P o (dP - rowsum(dP o P)) * scale (:1146-1156), dQ = dS K            */
/* (:1159-1168), dK = dS^T Q (:1171-1179).  The reference reads its fp16     */
/* inputs through a numeric cast of the bit pattern (main.mm:1100), which is */
/* a decoding bug (SURVEY.md section 4 defect 2); here the caller passes     */
/* properly decoded floats.  `causal` (keys j > i excluded, as               */
/* kernels.metal:1065-1079 masks them) is an extension: the reference's CPU  */
/* check is non-causal only.  Every sum runs in the reference's index order. */
/* ------------------------------------------------------------------------ */
void oracle_backward(const float *q, const float *k, const float *v, const float *dO,
                     float *dQ, float *dK, float *dV, int N, int D, float scale, int causal) {
  size_t nn = (size_t)N * (size_t)N;
  float *P = (float *)malloc(sizeof(float) * nn);
  float *dP = (float *)calloc(nn, sizeof(float));
  float *dS = (float *)calloc(nn, sizeof(float));
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < N; ++i) {
    int nk = causal ? i + 1 : N;
    float max_s = -INFINITY;
    for (int j = 0; j < N; ++j) P[(size_t)i * N + j] = 0.0f;
    for (int j = 0; j < nk; ++j) {
      float s = 0.0f;
      for (int d = 0; d < D; ++d) s += q[(long)i * D + d] * k[(long)j * D + d];
      s *= scale;
      P[(size_t)i * N + j] = s;
      if (s > max_s) max_s = s;
    }
    float sum_exp = 0.0f;
    for (int j = 0; j < nk; ++j) {
      P[(size_t)i * N + j] = expf(P[(size_t)i * N + j] - max_s);
      sum_exp += P[(size_t)i * N + j];
    }
    for (int j = 0; j < nk; ++j) P[(size_t)i * N + j] /= sum_exp;
  }
#pragma omp parallel for schedule(dynamic, 8)
  for (int j = 0; j < N; ++j)
    for (int d = 0; d < D; ++d) {
      float acc = 0.0f;
      for (int i = 0; i < N; ++i) acc += P[(size_t)i * N + j] * dO[(long)i * D + d];
      dV[(long)j * D + d] = acc;
    }
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < N; ++i) {
    int nk = causal ? i + 1 : N;
    for (int j = 0; j < nk; ++j) {
      float acc = 0.0f;
      for (int d = 0; d < D; ++d) acc += dO[(long)i * D + d] * v[(long)j * D + d];
      dP[(size_t)i * N + j] = acc;
    }
    float row_sum = 0.0f;
    for (int j = 0; j < N; ++j) row_sum += dP[(size_t)i * N + j] * P[(size_t)i * N + j];
    for (int j = 0; j < N; ++j)
      dS[(size_t)i * N + j] = P[(size_t)i * N + j] * (dP[(size_t)i * N + j] - row_sum) * scale;
  }
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < N; ++i)
    for (int d = 0; d < D; ++d) {
      float acc = 0.0f;
      for (int j = 0; j < N; ++j) acc += dS[(size_t)i * N + j] * k[(long)j * D + d];
      dQ[(long)i * D + d] = acc;
    }
#pragma omp parallel for schedule(dynamic, 8)
  for (int j = 0; j < N; ++j)
    for (int d = 0; d < D; ++d) {
      float acc = 0.0f;
      for (int i = 0; i < N; ++i) acc += dS[(size_t)i * N + j] * q[(long)i * D + d];
      dK[(long)j * D + d] = acc;
    }
  free(P);
  free(dP);
  free(dS);
}

/* fp64 backward (same formulas) for finite-difference and accuracy checks. */
void oracle_backward_f64(const float *q, const float *k, const float *v, const float *dO,
                         double *dQ, double *dK, double *dV, int N, int D, float scale,
                         int causal) {
  size_t nn = (size_t)N * (size_t)N;
  double *P = (double *)calloc(nn, sizeof(double));
  double *dS = (double *)calloc(nn, sizeof(double));
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < N; ++i) {
    int nk = causal ? i + 1 : N;
    double mx = -INFINITY;
    for (int j = 0; j < nk; ++j) {
      double s = 0.0;
      for (int d = 0; d < D; ++d) s += (double)q[(long)i * D + d] * (double)k[(long)j * D + d];
      s *= (double)scale;
      P[(size_t)i * N + j] = s;
      if (s > mx) mx = s;
    }
    double sum = 0.0;
    for (int j = 0; j < nk; ++j) { P[(size_t)i * N + j] = exp(P[(size_t)i * N + j] - mx); sum += P[(size_t)i * N + j]; }
    for (int j = 0; j < nk; ++j) P[(size_t)i * N + j] /= sum;
    double row_sum = 0.0;
    for (int j = 0; j < nk; ++j) {
      double acc = 0.0;
      for (int d = 0; d < D; ++d) acc += (double)dO[(long)i * D + d] * (double)v[(long)j * D + d];
      dS[(size_t)i * N + j] = acc;
      row_sum += acc * P[(size_t)i * N + j];
    }
    for (int j = 0; j < nk; ++j)
      dS[(size_t)i * N + j] = P[(size_t)i * N + j] * (dS[(size_t)i * N + j] - row_sum) * (double)scale;
  }
#pragma omp parallel for schedule(dynamic, 8)
  for (int j = 0; j < N; ++j)
    for (int d = 0; d < D; ++d) {
      double av = 0.0, ak = 0.0;
      for (int i = 0; i < N; ++i) {
        av += P[(size_t)i * N + j] * (double)dO[(long)i * D + d];
        ak += dS[(size_t)i * N + j] * (double)q[(long)i * D + d];
      }
      dV[(long)j * D + d] = av;
      dK[(long)j * D + d] = ak;
    }
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < N; ++i)
    for (int d = 0; d < D; ++d) {
      double acc = 0.0;
      for (int j = 0; j < N; ++j) acc += dS[(size_t)i * N + j] * (double)k[(long)j * D + d];
      dQ[(long)i * D + d] = acc;
    }
  free(P);
  free(dS);
}

/* ------------------------------------------------------------------------ */
/* Streaming backward for sizes where dense N x N buffers do not fit: same   */
/* formulas (main.mm:1091-1179), P recomputed from the row's max/sum, one    */
/* query row at a time; dK/dV accumulated per thread and reduced in thread   */
/* order.  Used only for large-N spot checks; not bit-identical to           */
/* oracle_backward (the j-sums of dK/dV are re-associated across threads).   */
/* ------------------------------------------------------------------------ */
void oracle_backward_streaming(const float *q, const float *k, const float *v, const float *dO,
                               float *dQ, float *dK, float *dV, int N, int D, float scale,
                               int causal) {
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  size_t nd = (size_t)N * (size_t)D;
  double *accK = (double *)calloc(nd * (size_t)nthreads, sizeof(double));
  double *accV = (double *)calloc(nd * (size_t)nthreads, sizeof(double));
#pragma omp parallel
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double *myK = accK + nd * (size_t)tid, *myV = accV + nd * (size_t)tid;
    float *p = (float *)malloc(sizeof(float) * (size_t)N);
    float *dp = (float *)malloc(sizeof(float) * (size_t)N);
#pragma omp for schedule(dynamic, 8)
    for (int i = 0; i < N; ++i) {
      int nk = causal ? i + 1 : N;
      float mx = -INFINITY;
      for (int j = 0; j < nk; ++j) {
        float s = 0.0f;
        for (int d = 0; d < D; ++d) s += q[(long)i * D + d] * k[(long)j * D + d];
        s *= scale;
        p[j] = s;
        if (s > mx) mx = s;
      }
      float sum = 0.0f;
      for (int j = 0; j < nk; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
      float row_sum = 0.0f;
      for (int j = 0; j < nk; ++j) {
        p[j] /= sum;
        float acc = 0.0f;
        for (int d = 0; d < D; ++d) acc += dO[(long)i * D + d] * v[(long)j * D + d];
        dp[j] = acc;
        row_sum += acc * p[j];
      }
      for (int d = 0; d < D; ++d) dQ[(long)i * D + d] = 0.0f;
      for (int j = 0; j < nk; ++j) {
        float ds = p[j] * (dp[j] - row_sum) * scale;
        for (int d = 0; d < D; ++d) {
          dQ[(long)i * D + d] += ds * k[(long)j * D + d];
          myK[(long)j * D + d] += (double)(ds * q[(long)i * D + d]);
          myV[(long)j * D + d] += (double)(p[j] * dO[(long)i * D + d]);
        }
      }
    }
    free(p);
    free(dp);
  }
  for (size_t e = 0; e < nd; ++e) {
    double sk = 0.0, sv = 0.0;
    for (int t = 0; t < nthreads; ++t) { sk += accK[nd * (size_t)t + e]; sv += accV[nd * (size_t)t + e]; }
    dK[e] = (float)sk;
    dV[e] = (float)sv;
  }
  free(accK);
  free(accV);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
