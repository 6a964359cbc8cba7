// fp64_worker.cuh -- EXPERIMENT, not used by the library (tools/int_peaks.cu only).
// Measured on B200: 16 FP64 warps alone 1.25e12 MAC/s; 12 IMAD + 4 FP64 warps on one tile 2.07e12 (< 2.39e12 of 16 IMAD
// warps): with equal sub-tiles the FP64 warps are the critical path, and an uneven split would add at most ~15 % to the
// matrix kernel for a second operand / accumulator format -- not adopted.
#pragma once
#include "../../pvw-rs_b200/csrc/mac_worker.cuh"

namespace pvw {

// ---------------------------------------------------------------------------------------------------------------------
// FP64-pipe worker.  The fma-heavy pipe saturates on the IMAD.WIDE carry chains of `Worker`; the FP64 pipe sits idle.
// A few warps per CTA therefore run the SAME multiply-accumulate as exact double-precision arithmetic: every residue
// (packed halves in shared memory) is re-split into pieces of 21 + 21 + 20 bits, converted with the 2^52 trick, and the
// 9 piece products (each < 2^42) are summed in five column accumulators.  A column holds at most 3 products per term, so
// it stays below 2^53 -- exactly representable -- for k <= 682 terms; the launcher uses this worker only for k <= 512.
// The epilogue converts the columns back to integers and reduces  sum_i col_i * 2^(21 i)  mod q like acck_reduce.
// ---------------------------------------------------------------------------------------------------------------------
struct DAcc {
  double c0, c1, c2, c3, c4;
};
struct DOp {
  double p0, p1, p2;
};
PVW_DEV DOp dsplit(u64 packed) {  // packed halves -> three exact doubles
  const u32 x0 = (u32)packed, x1 = (u32)(packed >> 32);
  const u32 q0 = x0 & 0x1fffffu, q1 = (x0 >> 21) | ((x1 & 0x7ffu) << 10), q2 = x1 >> 11;
  const double magic = 4503599627370496.0;  // 2^52: (0x43300000, p) is the double 2^52 + p
  DOp d;
  d.p0 = __hiloint2double(0x43300000, (int)q0) - magic;
  d.p1 = __hiloint2double(0x43300000, (int)q1) - magic;
  d.p2 = __hiloint2double(0x43300000, (int)q2) - magic;
  return d;
}
PVW_DEV void dacc_mac(DAcc& c, const DOp& a, const DOp& b) {
  c.c0 = fma(a.p0, b.p0, c.c0);
  c.c1 = fma(a.p0, b.p1, c.c1);
  c.c1 = fma(a.p1, b.p0, c.c1);
  c.c2 = fma(a.p0, b.p2, c.c2);
  c.c2 = fma(a.p1, b.p1, c.c2);
  c.c2 = fma(a.p2, b.p0, c.c2);
  c.c3 = fma(a.p1, b.p2, c.c3);
  c.c3 = fma(a.p2, b.p1, c.c3);
  c.c4 = fma(a.p2, b.p2, c.c4);
}
// canonical value mod q of  c0 + c1 2^21 + c2 2^42 + c3 2^63 + c4 2^84  (each column an exact integer < 2^53)
PVW_DEV u64 dacc_reduce(const DAcc& c, const LimbConst& lc) {
  const u64 v0 = (u64)c.c0, v1 = (u64)c.c1, v2 = (u64)c.c2, v3 = (u64)c.c3, v4 = (u64)c.c4;
  // 192-bit sum in three 64-bit words
  u64 w0 = v0, w1 = 0, w2 = 0;
  auto add_shifted = [&](u64 v, int sh) {  // += v << sh, sh in (0, 128)
    u64 a0, a1, a2;
    if (sh < 64) { a0 = v << sh; a1 = v >> (64 - sh); a2 = 0; }
    else { a0 = 0; a1 = v << (sh - 64); a2 = sh == 64 ? 0 : v >> (128 - sh); }
    const u64 s0 = w0 + a0, c0_ = s0 < a0;
    const u64 s1 = w1 + a1, c1a = s1 < a1, s1b = s1 + c0_, c1b = s1b < c0_;
    w0 = s0; w1 = s1b; w2 += a2 + c1a + c1b;
  };
  add_shifted(v1, 21);
  add_shifted(v2, 42);
  add_shifted(v3, 63);
  add_shifted(v4, 84);
  const u64 h = reduce128(reduce64(w2, lc), w1, lc);
  return reduce128(h, w0, lc);
}

template <int ELL, int TR, int TD, int GD, int KC, int THREADS_, bool DENSE_M = false>
struct DWorker {
  using C = TileCfg<ELL, TR, TD, GD, KC, THREADS_, DENSE_M>;
  int c, gr, gd;
  DAcc acc[TR][TD];
  __device__ __forceinline__ void init(int tid) {
    const int lane = tid & 31, w = tid >> 5;
    c = lane % ELL;
    const int g = w * (32 / ELL) + lane / ELL;
    gd = g % GD;
    gr = g / GD;
#pragma unroll
    for (int t = 0; t < TR; t++)
#pragma unroll
      for (int u = 0; u < TD; u++) acc[t][u] = DAcc{0.0, 0.0, 0.0, 0.0, 0.0};
  }
  __device__ __forceinline__ void chunk(const unsigned char* stage, int kc) {
    const unsigned char* ms = stage + (size_t)(gr * TR) * C::MROWB + c * 8;
    const unsigned char* vs = stage + C::MBYTES + (size_t)(gd * TD) * C::VROWB + c * 8;
#pragma unroll 2
    for (int jj = 0; jj < kc; jj++) {
      DOp b[TD];
#pragma unroll
      for (int u = 0; u < TD; u++) b[u] = dsplit(*reinterpret_cast<const u64*>(vs + u * C::VROWB + jj * ELL * 8));
#pragma unroll
      for (int t = 0; t < TR; t++) {
        const DOp a = dsplit(*reinterpret_cast<const u64*>(ms + t * C::MROWB + jj * ELL * 8));
#pragma unroll
        for (int u = 0; u < TD; u++) dacc_mac(acc[t][u], a, b[u]);
      }
    }
  }
};

}  // namespace pvw
