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
  u64 pad;      // always 0 (mac.cu)
  u64 c124;     // 2^124 mod q   (one-step reduction of the tensor-core product's 160-bit sums, imma.cu)
  u64 pad2;
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
// the same without the final correction: result in [0, 2q) for any a < 2^64 (Harvey's lazy butterflies; q < 2^62, so sums of
// up to four such values still fit 64 bits)
PVW_DEV u64 mulmod_shoup_lazy(u64 a, u64 w, u64 w_sh, u64 q) { return a * w - __umul64hi(a, w_sh) * q; }
// x in [0, 2q) -> [0, q);  x in [0, 4q) -> [0, 2q) with q2 = 2q
PVW_DEV u64 csub(u64 x, u64 q) { return x >= q ? x - q : x; }
// i64 -> canonical residue: ((x % q) + q) % q   (parameters.rs:437-452, Poly::from_coefficients)
PVW_DEV u64 reduce_i64(long long x, const LimbConst& c) {
  const u64 m = x >= 0 ? (u64)x : (u64)(-(x + 1)) + 1ull;  // |x| without overflow
  const u64 r = m < c.q ? m : reduce64(m, c);              // secrets / errors are tiny: no multiply on the common path
  return (x >= 0 || r == 0) ? r : c.q - r;
}

// ---------------------------------------------------------------------------------------------------------------
// [measured alternative, used only by tools/csrc/int_peaks.cu: 2.1e12 MAC/s in registers]
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

// ---------------------------------------------------------------------------------------------------------------
// [PRODUCTION accumulator of mac.cu: 2.8e12 MAC/s in registers]
// Three-multiply lazy accumulator (Karatsuba on 31-bit halves).  x = x1*2^31 + x0 with x0, x1 < 2^31 (x < 2^62) and
// xs = x0 + x1 < 2^32, so all three partial products are single 32x32->64 IMAD.WIDE.U32:
//     a*b = H*2^62 + (K - L - H)*2^31 + L,   L = a0*b0,  H = a1*b1,  K = as*bs.
// L, H and K are summed unreduced over the whole inner dimension in 96 bits each (k < 2^32 terms); the subtraction,
// the recombination and the single modular reduction happen once per output.  3 IMAD.WIDE.U32 (fma pipe) + 3 IADD3.X
// (alu pipe) per multiply-accumulate instead of 4 + 3.
// ---------------------------------------------------------------------------------------------------------------
struct SplitOp {
  u32 x0, x1, xs;
};
// 32-bit add pinned to the alu pipe: a plain `+` is often emitted as IMAD.IADD on the fma pipe, which the
// multiply-accumulate loop saturates.  A three-input add can only be an IADD3; `z` is a run-time zero the compiler
// cannot see through (LimbConst::pad).
PVW_DEV u32 add_alu(u32 a, u32 b, u32 z) { return a + b + z; }
PVW_DEV SplitOp split_op(u64 x, u32 z = 0) {
  SplitOp s;
  s.x0 = (u32)x & 0x7fffffffu;
  s.x1 = (u32)(x >> 31);
  s.xs = add_alu(s.x0, s.x1, z);
  return s;
}
struct AccK {
  u32 l0, l1, l2, h0, h1, h2, k0, k1, k2;
};
PVW_DEV void acck_zero(AccK& c) { c.l0 = c.l1 = c.l2 = c.h0 = c.h1 = c.h2 = c.k0 = c.k1 = c.k2 = 0; }
// PVW_SPLIT_L=1 (measured, slower in situ): form L = a0*b0 with a plain IMAD.WIDE (no addend) and add it into the 96-bit
// sum with three IADD3 on the alu pipe
#ifndef PVW_SPLIT_L
#define PVW_SPLIT_L 0
#endif
PVW_DEV void acck_mac(AccK& c, const SplitOp& a, const SplitOp& b) {
#if PVW_SPLIT_L
  {
    u32 plo, phi;
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(plo), "=r"(phi) : "r"(a.x0), "r"(b.x0));
    asm("add.cc.u32 %0, %0, %3;\n\t"
        "addc.cc.u32 %1, %1, %4;\n\t"
        "addc.u32 %2, %2, 0;\n\t"
        : "+r"(c.l0), "+r"(c.l1), "+r"(c.l2)
        : "r"(plo), "r"(phi));
  }
#else
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      : "+r"(c.l0), "+r"(c.l1), "+r"(c.l2)
      : "r"(a.x0), "r"(b.x0));
#endif
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      : "+r"(c.h0), "+r"(c.h1), "+r"(c.h2)
      : "r"(a.x1), "r"(b.x1));
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      : "+r"(c.k0), "+r"(c.k1), "+r"(c.k2)
      : "r"(a.xs), "r"(b.xs));
}
// canonical value of the accumulator mod q: recombine the three 96-bit sums into one 160-bit integer
//   T = L + (K - L - H) * 2^31 + H * 2^62      (K - L - H = sum of the cross products >= 0)
// with shifts and word adds, then two Barrett steps (10 wide multiplies instead of 34 for reducing L, H, K separately)
PVW_DEV u64 acck_reduce(const AccK& c, const LimbConst& lc) {
  // mid = K - L - H, 96 bits
  u32 m0, m1, m2;
  asm("sub.cc.u32 %0, %3, %6;\n\t"
      "subc.cc.u32 %1, %4, %7;\n\t"
      "subc.u32 %2, %5, %8;\n\t"
      "sub.cc.u32 %0, %0, %9;\n\t"
      "subc.cc.u32 %1, %1, %10;\n\t"
      "subc.u32 %2, %2, %11;\n\t"
      : "=&r"(m0), "=&r"(m1), "=&r"(m2)
      : "r"(c.k0), "r"(c.k1), "r"(c.k2), "r"(c.l0), "r"(c.l1), "r"(c.l2), "r"(c.h0), "r"(c.h1), "r"(c.h2));
  // mid << 31 -> words s0..s3 ; H << 62 -> words t1..t4 (62 = 32 + 30)
  const u32 s0 = m0 << 31, s1 = __funnelshift_l(m0, m1, 31), s2 = __funnelshift_l(m1, m2, 31), s3 = m2 >> 1;
  const u32 t1 = c.h0 << 30, t2 = __funnelshift_l(c.h0, c.h1, 30), t3 = __funnelshift_l(c.h1, c.h2, 30), t4 = c.h2 >> 2;
  u64 a = (u64)c.l0 + s0;
  const u32 T0 = (u32)a; a >>= 32;
  a += (u64)c.l1 + s1 + t1;
  const u32 T1 = (u32)a; a >>= 32;
  a += (u64)c.l2 + s2 + t2;
  const u32 T2 = (u32)a; a >>= 32;
  a += (u64)s3 + t3;
  const u32 T3 = (u32)a; a >>= 32;
  a += t4;  // < 2^32: T < 2^32 terms * 2^124
  const u64 lo = ((u64)T1 << 32) | T0, hi = ((u64)T3 << 32) | T2;
  const u64 h = reduce128(reduce64(a, lc), hi, lc);
  return reduce128(h, lo, lc);
}

// ---------------------------------------------------------------------------------------------------------------
// Packed operand form (production): x = x1*2^31 + x0 is stored as (x1 << 32) | x0, see kernels.cuh.
//
// [measured alternative, used only by tools/csrc/int_peaks.cu: 1.6e12 MAC/s -- ptxas re-associates the plain mad.wide chains
//  into IMAD.WIDE(RZ) + IADD3 + IMAD.X, so dropping the carries does not pay]
// Carry-free lazy accumulator: every partial product of 31-bit halves is < 2^62, so four consecutive terms are summed in
// plain 64-bit accumulators (no carry possible) before being folded into 96-bit sums:
//     a*b = H*2^62 + (M01 + M10)*2^31 + L,   L = a0*b0, M01 = a0*b1, M10 = a1*b0, H = a1*b1
// ---------------------------------------------------------------------------------------------------------------
PVW_DEV u64 pack_halves(u64 x) { return ((x >> 31) << 32) | (x & 0x7fffffffull); }       // x < 2^62
PVW_DEV u64 unpack_halves(u64 p) { return ((p >> 32) << 31) + (p & 0xffffffffull); }
PVW_DEV u64 mul_wide(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
PVW_DEV u64 mad_wide(u32 a, u32 b, u64 c) { u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
struct AccW {
  u32 l0, l1, l2, m0, m1, m2, h0, h1, h2;
};
PVW_DEV void accw_zero(AccW& c) { c.l0 = c.l1 = c.l2 = c.m0 = c.m1 = c.m2 = c.h0 = c.h1 = c.h2 = 0; }
PVW_DEV u32 lo32(u64 p) { u32 lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p)); return lo; }
PVW_DEV u32 hi32(u64 p) { u32 lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p)); return hi; }
PVW_DEV void fold96(u32& w0, u32& w1, u32& w2, u64 p) {
  asm("add.cc.u32 %0, %0, %3;\n\t"
      "addc.cc.u32 %1, %1, %4;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      : "+r"(w0), "+r"(w1), "+r"(w2)
      : "r"(lo32(p)), "r"(hi32(p)));
}
// N <= 4 consecutive terms: a[i], b[i] are packed halves
template <int N>
PVW_DEV void accw_mac(AccW& c, const u64 (&a)[N], const u64 (&b)[N]) {
  static_assert(N >= 1 && N <= 4, "at most four 62-bit products fit a 64-bit partial sum");
  u64 pl, ph, p01, p10;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const u32 a0 = (u32)a[i], a1 = (u32)(a[i] >> 32), b0 = (u32)b[i], b1 = (u32)(b[i] >> 32);
    if (i == 0) {
      pl = mul_wide(a0, b0); p01 = mul_wide(a0, b1); p10 = mul_wide(a1, b0); ph = mul_wide(a1, b1);
    } else {
      pl = mad_wide(a0, b0, pl); p01 = mad_wide(a0, b1, p01); p10 = mad_wide(a1, b0, p10); ph = mad_wide(a1, b1, ph);
    }
  }
  fold96(c.l0, c.l1, c.l2, pl);
  fold96(c.m0, c.m1, c.m2, p01);
  fold96(c.m0, c.m1, c.m2, p10);
  fold96(c.h0, c.h1, c.h2, ph);
}
// canonical value of the accumulator mod q
PVW_DEV u64 accw_reduce(const AccW& c, const LimbConst& lc) {
  const u64 L = reduce128(reduce64((u64)c.l2, lc), ((u64)c.l1 << 32) | c.l0, lc);
  const u64 M = reduce128(reduce64((u64)c.m2, lc), ((u64)c.m1 << 32) | c.m0, lc);
  const u64 H = reduce128(reduce64((u64)c.h2, lc), ((u64)c.h1 << 32) | c.h0, lc);
  const u64 p31 = reduce64(1ull << 31, lc), p62 = reduce64(1ull << 62, lc);
  return addmod(addmod(mulmod(H, p62, lc), mulmod(M, p31, lc), lc.q), L, lc.q);
}

// ---------------------------------------------------------------------------------------------------------------
// [measured alternative, used only by tools/csrc/int_peaks.cu: 2.0e12 MAC/s]
// Hybrid: Karatsuba with the two small products (L = a0*b0, H = a1*b1 < 2^62) summed carry-free four at a time and
// the cross term K = (a0+a1)*(b0+b1) < 2^64 accumulated with carry.  Packed operand word: (x1 << 32) | x0.
// ---------------------------------------------------------------------------------------------------------------
struct AccH {
  u32 l0, l1, l2, h0, h1, h2, k0, k1, k2;
};
PVW_DEV void acch_zero(AccH& c) { c.l0 = c.l1 = c.l2 = c.h0 = c.h1 = c.h2 = c.k0 = c.k1 = c.k2 = 0; }
template <int N>
PVW_DEV void acch_mac(AccH& c, const u64 (&a)[N], const u64 (&b)[N]) {
  static_assert(N >= 1 && N <= 4, "at most four 62-bit products fit a 64-bit partial sum");
  u64 pl, ph;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const u32 a0 = (u32)a[i], a1 = (u32)(a[i] >> 32), b0 = (u32)b[i], b1 = (u32)(b[i] >> 32);
    const u32 as = a0 + a1, bs = b0 + b1;
    if (i == 0) { pl = mul_wide(a0, b0); ph = mul_wide(a1, b1); }
    else { pl = mad_wide(a0, b0, pl); ph = mad_wide(a1, b1, ph); }
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;\n\t"
        : "+r"(c.k0), "+r"(c.k1), "+r"(c.k2)
        : "r"(as), "r"(bs));
  }
  fold96(c.l0, c.l1, c.l2, pl);
  fold96(c.h0, c.h1, c.h2, ph);
}

}  // namespace pvw
