// ntt_regs.cuh -- ell-point negacyclic NTT on a register array (ell in {8,16,32}); loops fully unrolled so that every
// index is a compile-time constant.  Forward: Cooley-Tukey, natural in / bit-reversed out, twiddle tw[m+i] =
// psi^brv(m+i).  Inverse: Gentleman-Sande with twi[h+i] = psi^-brv(h+i) and a final scale by ell^-1.
// Canonical residues in and out (fhe-math NttOperator::forward/backward produce fully reduced outputs).
#pragma once
#include "modarith.cuh"

namespace pvw {

template <int ELL>
PVW_DEV void ntt_forward_regs(u64 (&a)[ELL], const u64* tw, const u64* tw_sh, u64 q) {
  int t = ELL;
#pragma unroll
  for (int m = 1; m < ELL; m <<= 1) {
    t >>= 1;
#pragma unroll
    for (int i = 0; i < m; i++) {
      const u64 s = tw[m + i], s_sh = tw_sh[m + i];
      const int j1 = 2 * i * t;
#pragma unroll
      for (int j = j1; j < j1 + t; j++) {
        u64 u = a[j], v = mulmod_shoup(a[j + t], s, s_sh, q);
        a[j] = addmod(u, v, q);
        a[j + t] = submod(u, v, q);
      }
    }
  }
}

// the Gentleman-Sande passes without the final scale by ell^-1 (callers that multiply the result anyway fold it in)
template <int ELL>
PVW_DEV void ntt_inverse_unscaled_regs(u64 (&a)[ELL], const u64* twi, const u64* twi_sh, u64 q) {
  int t = 1;
#pragma unroll
  for (int m = ELL; m > 1; m >>= 1) {
    const int h = m >> 1;
    int j1 = 0;
#pragma unroll
    for (int i = 0; i < h; i++) {
      const u64 s = twi[h + i], s_sh = twi_sh[h + i];
#pragma unroll
      for (int j = j1; j < j1 + t; j++) {
        u64 u = a[j], v = a[j + t];
        a[j] = addmod(u, v, q);
        a[j + t] = mulmod_shoup(submod(u, v, q), s, s_sh, q);
      }
      j1 += 2 * t;
    }
    t <<= 1;
  }
}

// Lazy forms (Harvey): one conditional subtraction per butterfly instead of three.
// forward: canonical (or any < 4q) in, canonical out.  Invariant: every value stays below 4q < 2^64.
template <int ELL>
PVW_DEV void ntt_forward_lazy_regs(u64 (&a)[ELL], const u64* tw, const u64* tw_sh, u64 q) {
  const u64 q2 = 2 * q;
  int t = ELL;
#pragma unroll
  for (int m = 1; m < ELL; m <<= 1) {
    t >>= 1;
#pragma unroll
    for (int i = 0; i < m; i++) {
      const u64 s = tw[m + i], s_sh = tw_sh[m + i];
      const int j1 = 2 * i * t;
#pragma unroll
      for (int j = j1; j < j1 + t; j++) {
        const u64 u = csub(a[j], q2), v = mulmod_shoup_lazy(a[j + t], s, s_sh, q);   // u, v < 2q
        a[j] = u + v;                                                                // < 4q
        a[j + t] = u - v + q2;                                                       // < 4q
      }
    }
  }
#pragma unroll
  for (int j = 0; j < ELL; j++) a[j] = csub(csub(a[j], q2), q);
}
// inverse passes without the scale by ell^-1: canonical (or any < 2q) in, values in [0, 2q) out
template <int ELL>
PVW_DEV void ntt_inverse_unscaled_lazy_regs(u64 (&a)[ELL], const u64* twi, const u64* twi_sh, u64 q) {
  const u64 q2 = 2 * q;
  int t = 1;
#pragma unroll
  for (int m = ELL; m > 1; m >>= 1) {
    const int h = m >> 1;
    int j1 = 0;
#pragma unroll
    for (int i = 0; i < h; i++) {
      const u64 s = twi[h + i], s_sh = twi_sh[h + i];
#pragma unroll
      for (int j = j1; j < j1 + t; j++) {
        const u64 u = a[j], v = a[j + t];                                            // < 2q each
        a[j] = csub(u + v, q2);                                                      // < 2q
        a[j + t] = mulmod_shoup_lazy(u - v + q2, s, s_sh, q);                        // < 2q
      }
      j1 += 2 * t;
    }
    t <<= 1;
  }
}

template <int ELL>
PVW_DEV void ntt_inverse_regs(u64 (&a)[ELL], const u64* twi, const u64* twi_sh, u64 ninv, u64 ninv_sh, u64 q) {
  int t = 1;
#pragma unroll
  for (int m = ELL; m > 1; m >>= 1) {
    const int h = m >> 1;
    int j1 = 0;
#pragma unroll
    for (int i = 0; i < h; i++) {
      const u64 s = twi[h + i], s_sh = twi_sh[h + i];
#pragma unroll
      for (int j = j1; j < j1 + t; j++) {
        u64 u = a[j], v = a[j + t];
        a[j] = addmod(u, v, q);
        a[j + t] = mulmod_shoup(submod(u, v, q), s, s_sh, q);
      }
      j1 += 2 * t;
    }
    t <<= 1;
  }
#pragma unroll
  for (int j = 0; j < ELL; j++) a[j] = mulmod_shoup(a[j], ninv, ninv_sh, q);
}

}  // namespace pvw
