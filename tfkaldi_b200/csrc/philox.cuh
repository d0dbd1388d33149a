// Philox4x32-10 counter-based RNG, shared by the forward-dropout epilogue and (re-stated in numpy) the
// CPU oracle so both sides regenerate the SAME dropout mask from (seed, row, column).
// Semantics mirrored: tf.nn.dropout(x, keep) = x/keep * floor(keep + u), u ~ U[0,1)
// (reference: neuralNetworks/classifiers/activation.py:140-141).
// One Philox4x32-10 call yields 128 bits = EIGHT 16-bit draws: u = f * 2^-16 from field f, and the element is kept iff
// f >= thr, thr = ceil((1-keep) * 2^16), which is floor(keep + u) == 1 evaluated in exact integer arithmetic.
// Counter = (col >> 3, row, 0, 0), key = (seed lo, seed hi); output word i serves columns 8*(col>>3) + 2i (low half)
// and + 2i + 1 (high half).  (Eight decisions per call instead of four halves the integer work of the dropout passes,
// which were ALU-bound on it: profiles/r2b_ncu_full_summary_c4_small_kernels.txt.)
#pragma once
#include <cstdint>

namespace tfk {

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ inline uint32_t philox_mulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = philox_mulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = philox_mulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// keep-threshold on a 16-bit field
inline uint32_t dropout_threshold(double keep) {
  double t = (1.0 - keep) * 65536.0;
  uint32_t thr = static_cast<uint32_t>(t);
  if (static_cast<double>(thr) < t) ++thr;  // ceil
  return thr;
}

// the eight keep decisions of one call, bit k = column 8*(col>>3) + k
__host__ __device__ inline uint32_t dropout_keep_bits(const Philox4& r, uint32_t thr) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t bits = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 0; i < 4; ++i) {
    bits |= ((w[i] & 0xFFFFu) >= thr ? 1u : 0u) << (2 * i);
    bits |= ((w[i] >> 16) >= thr ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}

}  // namespace tfk
