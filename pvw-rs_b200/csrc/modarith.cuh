// modarith.cuh -- 64-bit modular arithmetic for sm_100a (device) shared by every kernel of the PVW hot path.
//
// Replaces fhe-math zq::Modulus (add/sub/mul/reduce) on the device.  All public results are canonical
// residues in [0, q), q < 2^62, so they are bit-identical to the reference's regardless of how they were computed.
#pragma once
#include <cstdint>

namespace pvw {

typedef unsigned long long u64;
typedef unsigned int u32;

// per-RNS-limb constants, built on the host (hostparams.hpp) and read through a device pointer
struct LimbConst {
  u64 q;        // the modulus
  u64 mu_hi;    // floor(2^128 / q) high word
  u64 mu_lo;    //                  low word
  u64 mu64;     // floor(2^64 / q)
  u64 r128;     // 2^128 mod q   (weight of the 5th accumulator word)
  u64 ninv;     // ell^-1 mod q            (inverse NTT scale)
  u64 ninv_sh;  // Shoup companion floor(ninv * 2^64 / q)
  u64 delta;    // Delta mod q             (decode, decryption.rs:61-75)
  u64 delta_sh;
  u64 qhinv;    // (Q/q)^-1 mod q          (CRT lift, fhe-math RnsContext)
  u64 qhinv_sh;
  u64 pad;
};

#define PVW_DEV __device__ __forceinline__

PVW_DEV u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
PVW_DEV u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
PVW_DEV u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }

// x mod q for any 64-bit x
PVW_DEV u64 reduce64(u64 x, const LimbConst& c) {
  u64 qh = __umul64hi(x, c.mu64);
  u64 r = x - qh * c.q;  // r < 3q < 2^64
  if (r >= 2 * c.q) r -= 2 * c.q;
  if (r >= c.q) r -= c.q;
  return r;
}
// (h * 2^64 + lo) mod q, requires h < q.  Barrett with mu = floor(2^128/q): the estimate misses the true quotient by
// at most 3, so the remainder is < 4q < 2^64.
PVW_DEV u64 reduce128(u64 h, u64 lo, const LimbConst& c) {
  u64 qh = h * c.mu_hi + __umul64hi(h, c.mu_lo) + __umul64hi(lo, c.mu_hi);
  u64 r = lo - qh * c.q;
  if (r >= 2 * c.q) r -= 2 * c.q;
  if (r >= c.q) r -= c.q;
  return r;
}
// a * b mod q for a, b < q   (zq::Modulus::mul)
PVW_DEV u64 mulmod(u64 a, u64 b, const LimbConst& c) { return reduce128(__umul64hi(a, b), a * b, c); }
// a * w mod q with w < q constant and w_sh = floor(w * 2^64 / q)   (zq::Modulus::mul_shoup), any a < 2^64
PVW_DEV u64 mulmod_shoup(u64 a, u64 w, u64 w_sh, u64 q) {
  u64 r = a * w - __umul64hi(a, w_sh) * q;  // in [0, 2q)
  return r >= q ? r - q : r;
}
// i64 -> canonical residue: ((x % q) + q) % q   (parameters.rs:437-452, Poly::from_coefficients)
PVW_DEV u64 reduce_i64(long long x, const LimbConst& c) {
  if (x >= 0) return reduce64((u64)x, c);
  u64 m = (u64)(-(x + 1)) + 1ull;
  u64 r = reduce64(m, c);
  return r ? c.q - r : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Lazy 160-bit accumulator for sum_j a_j * b_j with a_j, b_j < 2^62: products are added unreduced and one Barrett
// reduction runs after the whole k-term sum (k < 2^32).  The four 32x32 partial products go to an "even" column
// set (a0*b0 at bit 0, a1*b1 at bit 64; words e0..e4) and an "odd" one (a0*b1 + a1*b0 at bit 32; words o0..o2) so
// that every product lands on an aligned 64-bit register pair: ptxas turns each mad.lo.cc/madc.hi.cc pair into one
// IMAD.WIDE.U32 with carry-out and the addc into IADD3.X -- 4 IMAD.WIDE + ~1.5 IADD3.X per multiply-accumulate.
// ---------------------------------------------------------------------------------------------------------------
struct Acc160 {
  u32 e0, e1, e2, e3, e4, o0, o1, o2;
};
PVW_DEV void acc_zero(Acc160& c) { c.e0 = c.e1 = c.e2 = c.e3 = c.e4 = c.o0 = c.o1 = c.o2 = 0; }
PVW_DEV void acc_mac(Acc160& c, u64 a, u64 b) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t"
      "madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
      "madc.lo.cc.u32 %2, %7, %8, %2;\n\t"
      "madc.hi.cc.u32 %3, %7, %8, %3;\n\t"
      "addc.u32 %4, %4, 0;\n\t"
      : "+r"(c.e0), "+r"(c.e1), "+r"(c.e2), "+r"(c.e3), "+r"(c.e4)
      : "r"(a0), "r"(b0), "r"(a1), "r"(b1));
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      "mad.lo.cc.u32 %0, %5, %6, %0;\n\t"
      "madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      : "+r"(c.o0), "+r"(c.o1), "+r"(c.o2)
      : "r"(a0), "r"(b1), "r"(a1), "r"(b0));
}
// canonical value of the accumulator mod q
PVW_DEV u64 acc_reduce(const Acc160& c, const LimbConst& lc) {
  // fold the odd columns in: total = e + (o << 32)
  u32 w0 = c.e0, w1, w2, w3, w4;
  asm("add.cc.u32 %0, %4, %7;\n\t"
      "addc.cc.u32 %1, %5, %8;\n\t"
      "addc.cc.u32 %2, %6, %9;\n\t"
      "addc.u32 %3, %10, 0;\n\t"
      : "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4)
      : "r"(c.e1), "r"(c.e2), "r"(c.e3), "r"(c.o0), "r"(c.o1), "r"(c.o2), "r"(c.e4));
  u64 lo = ((u64)w1 << 32) | w0, hi = ((u64)w3 << 32) | w2;
  u64 h = reduce64(hi, lc);
  u64 t = reduce128(h, lo, lc);
  u64 p_hi = __umul64hi((u64)w4, lc.r128), p_lo = (u64)w4 * lc.r128;  // w4 * r128 < 2^94: p_hi < 2^30 <= q? not for tiny q
  u64 u = reduce128(reduce64(p_hi, lc), p_lo, lc);
  return addmod(t, u, lc.q);
}

}  // namespace pvw
