// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Function shells around the reference's own CPU verifier loops.  The loop
// bodies are NOT in this repository: oracle/build_ref.sh cuts them, unmodified,
// out of /root/reference/main.mm into oracle/_ref/*.inc (git-ignored) and this
// file #includes them.  Only the declarations the loops expect from their
// surroundings in main() are supplied here (names as in main.mm).
//
//   ref_init_random.inc  = main.mm:24-30    (initRandom)
//   ref_forward.inc      = main.mm:128-159  (non-causal forward, i->d->j->k)
//   ref_causal.inc       = main.mm:550-578  (causal forward)
//   ref_backward.inc     = main.mm:1092-1179 (P, dV, dP, dS, dQ, dK)
//
// Two adaptations, both outside the lifted lines:
//  * `exp` is unqualified in main.mm.  Under Apple's libc++ that resolves to the
//    float overload; under glibc it would silently pick exp(double).  A
//    using-declaration in each shell restores the float overload.
//  * main.mm:1100 reads fp16 storage as `(float)((__fp16)q_h_ptr[i])` with
//    q_h_ptr a uint16_t*, i.e. it converts the bit pattern numerically
//    (SURVEY.md section 4 defect 2).  Here q_h_ptr/do_h_ptr point at HalfBits,
//    whose conversion operator decodes the bits, so the same source line reads
//    the value the GPU kernels see.  ref_backward_buggy() keeps the reference's
//    literal behaviour for the record.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

typedef _Float16 __fp16;

#include "_ref/ref_init_random.inc"

struct HalfBits {
  uint16_t bits;
  operator __fp16() const {
    __fp16 h;
    std::memcpy(&h, &bits, 2);
    return h;
  }
};

extern "C" {

void ref_init_random(float *data, int size) { initRandom(data, size); }

void ref_forward(const float *q_ptr, const float *k_ptr, const float *v_ptr, float *out, int N,
                 int D, float SCALE) {
  using std::exp;
  std::vector<float> O_cpu((size_t)N * D);
#include "_ref/ref_forward.inc"
  std::memcpy(out, O_cpu.data(), sizeof(float) * (size_t)N * D);
}

void ref_forward_causal(const float *qc_f, const float *kc_f, const float *vc_f, float *out,
                        int N_causal, int D, float SCALE) {
  using std::exp;
#include "_ref/ref_causal.inc"
  std::memcpy(out, O_ref.data(), sizeof(float) * (size_t)N_causal * D);
}

// Q (= K = V, as main.mm:966-967 copies them) and dO as fp16 bit patterns.
void ref_backward(const uint16_t *q_bits, const uint16_t *do_bits, float *dQ_out, float *dK_out,
                  float *dV_out, int curr_n, int D, float SCALE) {
  using std::exp;
  const HalfBits *q_h_ptr = reinterpret_cast<const HalfBits *>(q_bits);
  const HalfBits *do_h_ptr = reinterpret_cast<const HalfBits *>(do_bits);
#include "_ref/ref_backward.inc"
  std::memcpy(dQ_out, dQ_cpu.data(), sizeof(float) * (size_t)curr_n * D);
  std::memcpy(dK_out, dK_cpu.data(), sizeof(float) * (size_t)curr_n * D);
  std::memcpy(dV_out, dV_cpu.data(), sizeof(float) * (size_t)curr_n * D);
}

// The same lines with the reference's own types: uint16_t* -> numeric cast.
void ref_backward_buggy(const uint16_t *q_h_ptr, const uint16_t *do_h_ptr, float *dQ_out,
                        int curr_n, int D, float SCALE) {
  using std::exp;
#include "_ref/ref_backward.inc"
  std::memcpy(dQ_out, dQ_cpu.data(), sizeof(float) * (size_t)curr_n * D);
  (void)dK_cpu;
  (void)dV_cpu;
}

}  // extern "C"
