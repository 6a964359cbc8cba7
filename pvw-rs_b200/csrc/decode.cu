// decode.cu -- K4: the l-redundant PVW decoding of the noisy message zhat = <s,c1> - c2 (src/crypto/decryption.rs:10-58
// with its helpers :61-247), restated so that every step is either per-limb modular arithmetic or arithmetic on one
// multi-precision integer.  The reference moves between representations ~4l times per share (clone,
// inverse NTT, CRT lift of all l coefficients, BigInt op, re-encode, forward NTT); every one of those "constant
// polynomial" operations is arithmetic on a scalar mod Q, so the same integers are obtained by:
//
//   (1) decode_rns  : one inverse NTT per limb, then IN RNS  tmp_i = z_i*D - z_{i+1}  (:19-27),
//                     last = Horner(tmp, D) (:30-33), w = -z_0 (:51-52); pre-multiplied by (Q/q_j)^-1 for the lift;
//   (2) crt_lift    : CRT lift of the l+1 values {tmp_0..tmp_{l-2}, last, w} to integers in [0,Q)
//                     (Vec<BigUint>::from(&Poly), :118,:213 -- fhe-math RnsContext::lift);
//   (3) decode_tail : centred remainder of `last` by D^(l-1) (:154-178), the back-substitution
//                     noise_i = round((noise_{i+1} - tmp_i)/D) with truncated division (:44-48,:180-207),
//                     plaintext = -z_0 - noise_0 (:51-53) and the u64 conversion rules (:226-247).
//
// Results are the reference's for every input, including shares whose noise is too large to decode correctly.
#include <algorithm>

#include "kernels.cuh"
#include "ntt_regs.cuh"

// Minimum resident CTAs per SM asked of ptxas for the lift and tail kernels.  Both are latency-bound at the occupancy their
// natural register counts allow (ncu: 98 registers -> 24 % of the warp slots, 154 -> 18 %); capping them at 80 / 128
// registers costs a few spilled words and buys 1.62 -> 1.46 ms and 0.93 -> 0.80 ms per bench step.  (8, 5) was measured slower.
#ifndef PVW_LIFT_MINB
#define PVW_LIFT_MINB 6
#endif
#ifndef PVW_TAIL_MINB
#define PVW_TAIL_MINB 4
#endif
// resident CTAs per SM asked for the fused decode kernel at l = 8: 4 (128 registers, ~400 bytes spilled) measured 1.40 ms per
// bench step against 1.52 ms for 3 (168 registers, no spills) -- the kernel is latency bound and wants the warps
#ifndef PVW_FUSED_P2_MINB
#define PVW_FUSED_P2_MINB 6   // CTAs of 128 threads per SM the claim-check launch of the two-launch form is compiled for (80 registers)
#endif
#ifndef PVW_FUSED_MINB8
#define PVW_FUSED_MINB8 4
#endif

namespace pvw {

// ---------------------------------------------------------------------------------------------------------------
// (1) thread = (share, limb)
// ---------------------------------------------------------------------------------------------------------------
template <int ELL>
__global__ void __launch_bounds__(128) decode_rns_kernel(const u64* __restrict__ z, size_t z_ls, size_t z_ds, uint32_t Pc, uint64_t S,
                                                         u64* __restrict__ y, const LimbConst* __restrict__ lcs,
                                                         const u64* __restrict__ twi, const u64* __restrict__ twi_sh, size_t z_cs,
                                                         const DecodeSub sub, const u64* __restrict__ dec_c, const FallbackList fb) {
  __shared__ u64 s_tw[ELL], s_tw_sh[ELL];
  const uint32_t limb = blockIdx.y;
  if (threadIdx.x < ELL) {
    s_tw[threadIdx.x] = twi[(size_t)limb * ELL + threadIdx.x];
    s_tw_sh[threadIdx.x] = twi_sh[(size_t)limb * ELL + threadIdx.x];
  }
  __syncthreads();
  const LimbConst lc = lcs[limb];
  const uint64_t limit = fb.count ? *fb.count : S;   // list-driven: only the shares the fused fast path handed over
  for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < limit; it += (uint64_t)gridDim.x * blockDim.x) {
  const uint64_t s = fb.count ? fb.list[it] : it;
  const uint64_t d = s / Pc, p = s % Pc;
  u64 a[ELL];
  if (z_cs) {  // slot-major: consecutive threads read consecutive words of every slot plane
    const u64* src = z + d * z_ds + (size_t)limb * z_ls + p;
#pragma unroll
    for (int t = 0; t < ELL; t++) a[t] = src[(size_t)t * z_cs];
  } else {
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(z + d * z_ds + (size_t)limb * z_ls + p * ELL);
#pragma unroll
    for (int t = 0; t < ELL / 2; t++) {
      ulonglong2 v = src[t];
      a[2 * t] = v.x;
      a[2 * t + 1] = v.y;
    }
  }
  if (sub.S) {  // noisy = <s, c1> - c2[party]   (decryption.rs:270-274)
    const uint32_t sd = sub.dmap ? sub.dmap[d] : (uint32_t)d, srow = sub.rowmap ? sub.rowmap[p] : (uint32_t)p;
    const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(sub.S + (size_t)sd * sub.S_ds + (size_t)limb * sub.S_ls + (size_t)srow * ELL);
#pragma unroll
    for (int t = 0; t < ELL / 2; t++) {
      const ulonglong2 v = sp[t];
      a[2 * t] = submod(a[2 * t], v.x, lc.q);
      a[2 * t + 1] = submod(a[2 * t + 1], v.y, lc.q);
    }
  }
  // a' = ell * INTT(z): the scale ell^-1 is folded into the two multipliers c1 = Delta * ell^-1, c2 = ell^-1.  The small values
  // (tmp_i, -z_0) are stored as plain residues -- the short lift verifies its candidate against them without a multiply;
  // only `last`, which always takes the full CRT lift, is pre-multiplied by (Q/q)^-1.
  ntt_inverse_unscaled_regs<ELL>(a, s_tw, s_tw_sh, lc.q);
  const u64 q = lc.q;
  const u64 c1 = dec_c[4 * limb], c1_sh = dec_c[4 * limb + 1], c2 = dec_c[4 * limb + 2], c2_sh = dec_c[4 * limb + 3];
  u64* yo = y + ((size_t)limb * (ELL + 1)) * S + s;
  u64 last = 0, p2 = mulmod_shoup(a[0], c2, c2_sh, q);                                      // z_0
  yo[(size_t)ELL * S] = negmod(p2, q);                                                       // z_0 * (-1), decryption.rs:52
#pragma unroll
  for (int i = 0; i < ELL - 1; i++) {
    p2 = mulmod_shoup(a[i + 1], c2, c2_sh, q);
    const u64 tmp = submod(mulmod_shoup(a[i], c1, c1_sh, q), p2, q);                          // z_i * Delta - z_{i+1}, decryption.rs:25
    last = (i == 0) ? tmp : addmod(mulmod_shoup(last, lc.delta, lc.delta_sh, q), tmp, q);    // Horner, decryption.rs:30-33
    yo[(size_t)i * S] = tmp;
  }
  yo[(size_t)(ELL - 1) * S] = mulmod_shoup(last, lc.qhinv, lc.qhinv_sh, q);
  }
}

// the same step for any power-of-two ring degree up to 256 (run-time loops, coefficients in local memory): correct, not tuned
constexpr int GEN_MAX_ELL = 256;
__global__ void __launch_bounds__(64) decode_rns_generic_kernel(const u64* __restrict__ z, size_t z_ls, size_t z_ds, uint32_t Pc, uint64_t S,
                                                                u64* __restrict__ y, const LimbConst* __restrict__ lcs, const u64* __restrict__ twi,
                                                                const u64* __restrict__ twi_sh, size_t z_cs, const DecodeSub sub,
                                                                const u64* __restrict__ dec_c, const uint32_t ell, const FallbackList fb) {
  const uint32_t limb = blockIdx.y;
  const LimbConst lc = lcs[limb];
  const u64 q = lc.q;
  const uint64_t limit = fb.count ? *fb.count : S;
  for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < limit; it += (uint64_t)gridDim.x * blockDim.x) {
  const uint64_t s = fb.count ? fb.list[it] : it;
  const uint64_t d = s / Pc, p = s % Pc;
  u64 a[GEN_MAX_ELL];
  for (uint32_t t = 0; t < ell; t++)
    a[t] = z_cs ? z[d * z_ds + (size_t)limb * z_ls + (size_t)t * z_cs + p] : z[d * z_ds + (size_t)limb * z_ls + p * ell + t];
  if (sub.S) {
    const uint32_t sd = sub.dmap ? sub.dmap[d] : (uint32_t)d, srow = sub.rowmap ? sub.rowmap[p] : (uint32_t)p;
    const u64* sp = sub.S + (size_t)sd * sub.S_ds + (size_t)limb * sub.S_ls + (size_t)srow * ell;
    for (uint32_t t = 0; t < ell; t++) a[t] = submod(a[t], sp[t], q);
  }
  const u64* w = twi + (size_t)limb * ell;
  const u64* w_sh = twi_sh + (size_t)limb * ell;
  for (uint32_t m = ell, t = 1; m > 1; m >>= 1, t <<= 1) {   // Gentleman-Sande passes, unscaled (ntt_regs.cuh)
    const uint32_t h = m >> 1;
    for (uint32_t i = 0, j1 = 0; i < h; i++, j1 += 2 * t) {
      const u64 sw = w[h + i], sw_sh = w_sh[h + i];
      for (uint32_t j = j1; j < j1 + t; j++) {
        const u64 u = a[j], v = a[j + t];
        a[j] = addmod(u, v, q);
        a[j + t] = mulmod_shoup(submod(u, v, q), sw, sw_sh, q);
      }
    }
  }
  const u64 c1 = dec_c[4 * limb], c1_sh = dec_c[4 * limb + 1], c2 = dec_c[4 * limb + 2], c2_sh = dec_c[4 * limb + 3];
  u64* yo = y + ((size_t)limb * (ell + 1)) * S + s;
  u64 last = 0, p2 = mulmod_shoup(a[0], c2, c2_sh, q);
  yo[(size_t)ell * S] = negmod(p2, q);
  for (uint32_t i = 0; i + 1 < ell; i++) {
    p2 = mulmod_shoup(a[i + 1], c2, c2_sh, q);
    const u64 tmp = submod(mulmod_shoup(a[i], c1, c1_sh, q), p2, q);
    last = (i == 0) ? tmp : addmod(mulmod_shoup(last, lc.delta, lc.delta_sh, q), tmp, q);
    yo[(size_t)i * S] = tmp;
  }
  yo[(size_t)(ell - 1) * S] = mulmod_shoup(last, lc.qhinv, lc.qhinv_sh, q);
  }
}

// blocks of a launch over S shares: all of them, or -- list-driven (the fused fast path ran first and the list is usually
// empty) -- a few per SM that walk the list
static unsigned share_blocks(uint64_t S, unsigned threads, const FallbackList& fb) {
  const uint64_t full = (S + threads - 1) / threads;
  return (unsigned)(fb.count ? std::min<uint64_t>(full, 148u * 8u) : full);
}

bool launch_decode_rns(const DevTables& T, const u64* z, size_t z_ls, size_t z_ds, uint32_t Pc, uint32_t D, u64* y, cudaStream_t st, size_t z_cs,
                       const DecodeSub* sub, const FallbackList* fbl) {
  const uint64_t S = (uint64_t)Pc * D;
  if (S == 0) return true;
  if ((S + 63) / 64 >= (1ull << 31) || T.L > 65535u) return false;
  const FallbackList fb = fbl ? *fbl : FallbackList{nullptr, nullptr};
  dim3 grid(share_blocks(S, 128, fb), T.L);
  const DecodeSub sb = sub ? *sub : DecodeSub{nullptr, 0, 0, nullptr, nullptr};
  switch (T.ell) {
    case 8: decode_rns_kernel<8><<<grid, 128, 0, st>>>(z, z_ls, z_ds, Pc, S, y, T.lc, T.twi, T.twi_sh, z_cs, sb, T.dec_c, fb); break;
    case 16: decode_rns_kernel<16><<<grid, 128, 0, st>>>(z, z_ls, z_ds, Pc, S, y, T.lc, T.twi, T.twi_sh, z_cs, sb, T.dec_c, fb); break;
    case 32: decode_rns_kernel<32><<<grid, 128, 0, st>>>(z, z_ls, z_ds, Pc, S, y, T.lc, T.twi, T.twi_sh, z_cs, sb, T.dec_c, fb); break;
    default:
      if (T.ell > (uint32_t)GEN_MAX_ELL) return false;
      decode_rns_generic_kernel<<<dim3(share_blocks(S, 64, fb), T.L), 64, 0, st>>>(z, z_ls, z_ds, Pc, S, y, T.lc, T.twi, T.twi_sh, z_cs, sb, T.dec_c, T.ell, fb);
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// (2) thread = (share, value i): acc = sum_j y_j * (Q/q_j) (< L*Q) as 32-bit words in registers, then mod Q by
//     conditional subtraction of Q << b.  Row-by-word products run as carry chains (mad.lo.cc / madc.hi.cc ->
//     IMAD.WIDE with carry, even and odd columns separately so that every product lands on an aligned word pair);
//     the Q/q_j rows live in shared memory (uniform across the block).
// ---------------------------------------------------------------------------------------------------------------
template <int W, int OFF, int N>
PVW_DEV void mul_row_acc(u32 (&acc)[N], u32 m, const u32 (&q)[W]) {
  static_assert(W % 2 == 0 && OFF + W + 1 < N, "accumulator too short");
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[OFF]), "+r"(acc[OFF + 1]) : "r"(m), "r"(q[0]));
#pragma unroll
  for (int i = 2; i < W; i += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[OFF + i]), "+r"(acc[OFF + i + 1]) : "r"(m), "r"(q[i]));
  asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[OFF + W]));
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[OFF + 1]), "+r"(acc[OFF + 2]) : "r"(m), "r"(q[1]));
#pragma unroll
  for (int i = 3; i < W; i += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[OFF + i]), "+r"(acc[OFF + i + 1]) : "r"(m), "r"(q[i]));
  asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[OFF + W + 1]));
}

// Short lift (fast path of the CRT lift).  The values a decodable share produces (tmp_i ~ Delta * noise, w = m + noise)
// are far smaller than Q, so they are determined by a sub-basis q_0..q_{Ls-1} (product Q_s): lift there, centre, and
// VERIFY the candidate x against every remaining residue.  If all L residues agree, x = v (mod Q) with |x| <= Q_s/2 < Q/2,
// i.e. x is exactly the centred value of the full lift -- the result is bit-identical by CRT uniqueness; any mismatch
// (undecodable / adversarial inputs) falls through to the full lift below.
template <int SW>
PVW_DEV bool short_lift(const u64* __restrict__ y, size_t ystride, uint32_t L, const DevTables& T, const u64* __restrict__ vc,
                        u64 (&mag)[4], bool& neg) {
  u64 acc[SW + 1];
#pragma unroll
  for (int w = 0; w <= SW; w++) acc[w] = 0;
  const uint32_t Ls = T.shortL;
  for (uint32_t j = 0; j < Ls; j++) {
    const u64 t = mulmod_shoup(y[(size_t)j * ystride], T.sh_c[j], T.sh_c_sh[j], T.lc[j].q);
    u64 carry = 0;
#pragma unroll
    for (int w = 0; w < SW; w++) {
      const u64 qw = T.sh_qhat[(size_t)j * SW + w];
      const u64 lo = t * qw, hi = __umul64hi(t, qw);
      u64 x = acc[w] + carry;
      const u64 c1 = x < carry;
      x += lo;
      const u64 c2 = x < lo;
      acc[w] = x;
      carry = hi + c1 + c2;
    }
    acc[SW] += carry;
  }
  for (uint32_t r = 1; r < Ls; r++) {  // acc < Ls * Q_s
    bool ge = acc[SW] != 0;
    if (!ge) {
      ge = true;
#pragma unroll
      for (int w = SW - 1; w >= 0; w--) {
        const u64 qw = T.sh_Q[w];
        if (acc[w] != qw) { ge = acc[w] > qw; break; }
      }
    }
    if (ge) {
      u64 borrow = 0;
#pragma unroll
      for (int w = 0; w < SW; w++) {
        const u64 qw = T.sh_Q[w], d1 = acc[w] - qw, b1 = acc[w] < qw, d2 = d1 - borrow, b2 = d1 < borrow;
        acc[w] = d2;
        borrow = b1 | b2;
      }
      acc[SW] -= borrow;
    }
  }
  neg = false;
#pragma unroll
  for (int w = SW - 1; w >= 0; w--) {
    const u64 hw = T.sh_halfQ[w];
    if (acc[w] != hw) { neg = acc[w] > hw; break; }
  }
  if (neg) {
    u64 borrow = 0;
#pragma unroll
    for (int w = 0; w < SW; w++) {
      const u64 qw = T.sh_Q[w], d1 = qw - acc[w], b1 = qw < acc[w], d2 = d1 - borrow, b2 = d1 < borrow;
      acc[w] = d2;
      borrow = b1 | b2;
    }
  }
#pragma unroll
  for (int w = 0; w < 4; w++) mag[w] = w < SW ? acc[w] : 0;
  bool ok = true;
#pragma unroll 4
  for (uint32_t j = Ls; j < L; j++) {
    const u64* cj = vc + (size_t)j * 10;  // shared memory: q, floor(2^64/q), (Q/q) mod q + Shoup, 2^64 / 2^128 / 2^192 mod q + Shoup
    const u64 q = cj[0];
    u64 r = mag[0] - __umul64hi(mag[0], cj[1]) * q;  // < 3q
    r = r >= 2 * q ? r - 2 * q : r;
    r = r >= q ? r - q : r;
#pragma unroll
    for (int w = 1; w < SW; w++) r += mulmod_shoup(mag[w], cj[2 + 2 * w], cj[3 + 2 * w], q);  // < SW * q < 2^64
#pragma unroll
    for (int w = 1; w < SW; w++) r = r >= q ? r - q : r;
    if (neg) r = r ? q - r : 0;
    const u64 yj = y[(size_t)j * ystride];                                                  // the plain residue decode_rns stored
    ok = ok & (r == yj);
  }
  return ok;
}

template <int NWT>
__global__ void __launch_bounds__(128, PVW_LIFT_MINB) crt_lift_kernel(const u64* __restrict__ y, u64* __restrict__ X, uint64_t S, uint32_t L, uint32_t ellp1,
                                                       uint32_t NW, const u64* __restrict__ qhat, const u64* __restrict__ Qsh, uint32_t LB, const DevTables T,
                                                       const FallbackList fb) {
  constexpr int W = 2 * NWT;   // 32-bit words of one Q/q_j row
  constexpr int N = W + 3;     // accumulator words: the sum is < L*Q < 2^(32 W + 7)
  extern __shared__ __align__(16) u32 s_q[];  // [L][W] then [LB][W + 2]
  {
    const u32* g = reinterpret_cast<const u32*>(qhat);
    for (uint32_t i = threadIdx.x; i < L * W; i += blockDim.x) s_q[i] = g[i];
    const u32* gs = reinterpret_cast<const u32*>(Qsh);
    for (uint32_t i = threadIdx.x; i < LB * (W + 2); i += blockDim.x) s_q[L * W + i] = gs[i];
  }
  u64* s_vc = reinterpret_cast<u64*>(s_q + (((size_t)L * W + (size_t)LB * (W + 2) + 3) & ~(size_t)3));  // [L][10], 16-byte aligned
  if (T.lift_fast) {
    for (uint32_t j = threadIdx.x; j < L; j += blockDim.x) {
      u64* cj = s_vc + (size_t)j * 10;
      cj[0] = T.lc[j].q; cj[1] = T.lc[j].mu64; cj[2] = T.sh_v[j]; cj[3] = T.sh_v_sh[j];
      for (int t = 0; t < 3; t++) { cj[4 + 2 * t] = T.sh_r[(size_t)j * 3 + t]; cj[5 + 2 * t] = T.sh_r_sh[(size_t)j * 3 + t]; }
    }
  }
  __syncthreads();
  const uint32_t i = blockIdx.y;
  const uint64_t limit = fb.count ? *fb.count : S;
  for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < limit; it += (uint64_t)gridDim.x * blockDim.x) {
  const uint64_t s = fb.count ? fb.list[it] : it;
  if (T.lift_fast && i + 2 != ellp1) {  // every value except `last` (index l-1) is small for a decodable share
    u64 mag[4];
    bool neg = false, ok = false;
    const u64* yb = y + (size_t)i * S + s;
    const size_t ystride = (size_t)ellp1 * S;
    switch (T.shortSW) {
      case 1: ok = short_lift<1>(yb, ystride, L, T, s_vc, mag, neg); break;
      case 2: ok = short_lift<2>(yb, ystride, L, T, s_vc, mag, neg); break;
      case 3: ok = short_lift<3>(yb, ystride, L, T, s_vc, mag, neg); break;
      case 4: ok = short_lift<4>(yb, ystride, L, T, s_vc, mag, neg); break;
    }
    if (ok) {
      u64* xo = X + ((size_t)i * NW) * S + s;
      const bool zero = (mag[0] | mag[1] | mag[2] | mag[3]) == 0;
      if (!neg || zero) {
        for (uint32_t w = 0; w < NW; w++) xo[(size_t)w * S] = w < 4 ? mag[w] : 0;
      } else {  // Q - |x|
        u64 borrow = 0;
        for (uint32_t w = 0; w < NW; w++) {
          const u64 qw = T.Qw[w], mw = w < 4 ? mag[w] : 0, d1 = qw - mw, b1 = qw < mw, d2 = d1 - borrow, b2 = d1 < borrow;
          xo[(size_t)w * S] = d2;
          borrow = b1 | b2;
        }
      }
      continue;
    }
  }
  u32 acc[N];
#pragma unroll
  for (int w = 0; w < N; w++) acc[w] = 0;
  for (uint32_t j = 0; j < L; j++) {
    u64 yv = y[((size_t)j * ellp1 + i) * S + s];
    if (i + 2 != ellp1) yv = mulmod_shoup(yv, T.lc[j].qhinv, T.lc[j].qhinv_sh, T.lc[j].q);   // every value but `last` arrives as a plain residue
    u32 q[W];
    const uint2* row = reinterpret_cast<const uint2*>(s_q + (size_t)j * W);
#pragma unroll
    for (int w = 0; w < NWT; w++) { const uint2 v = row[w]; q[2 * w] = v.x; q[2 * w + 1] = v.y; }
    mul_row_acc<W, 0>(acc, (u32)yv, q);
    mul_row_acc<W, 1>(acc, (u32)(yv >> 32), q);
  }
  for (int b = (int)LB - 1; b >= 0; b--) {
    const u32* qs = s_q + (size_t)L * W + (size_t)b * (W + 2);
    // borrow of acc - (Q << b) without storing the difference
    u32 t, borrow;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(t) : "r"(acc[0]), "r"(qs[0]));
#pragma unroll
    for (int w = 1; w < W + 2; w++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(t) : "r"(acc[w]), "r"(qs[w]));
    asm volatile("subc.u32 %0, 0, 0;" : "=r"(borrow));
    if (borrow == 0) {  // acc >= Q << b
      asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(acc[0]) : "r"(qs[0]));
#pragma unroll
      for (int w = 1; w < W + 2; w++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(acc[w]) : "r"(qs[w]));
    }
  }
  u64* xo = X + ((size_t)i * NW) * S + s;
#pragma unroll
  for (int w = 0; w < NWT; w++)
    if (w < (int)NW) xo[(size_t)w * S] = ((u64)acc[2 * w + 1] << 32) | acc[2 * w];
  }
}

void launch_crt_lift(const DevTables& T, const u64* y, u64* X, uint64_t S, cudaStream_t st, const FallbackList* fbl) {
  if (S == 0) return;
  const FallbackList fb = fbl ? *fbl : FallbackList{nullptr, nullptr};
  dim3 grid(share_blocks(S, 128, fb), T.ell + 1);
  const size_t smem = ((size_t)T.L * 2 * T.NWT + (size_t)T.LB * (2 * T.NWT + 2) + 4) * 4 + (size_t)T.L * 10 * 8;
#define PVW_LIFT_CASE(N)                                                                                              \
  case N: {                                                                                                           \
    auto kern = crt_lift_kernel<N>;                                                                                   \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);                              \
    kern<<<grid, 128, smem, st>>>(y, X, S, T.L, T.ell + 1, T.NW, T.qhat, T.Qsh, T.LB, T, fb);                         \
    break;                                                                                                            \
  }
  switch (T.NWT) {
    PVW_LIFT_CASE(2)
    PVW_LIFT_CASE(4)
    PVW_LIFT_CASE(8)
    PVW_LIFT_CASE(17)
    PVW_LIFT_CASE(33)
    PVW_LIFT_CASE(64)
  }
#undef PVW_LIFT_CASE
}

// ---------------------------------------------------------------------------------------------------------------
// (3) thread = share; multi-precision integers as little-endian u64 arrays in local memory (runtime length).
// ---------------------------------------------------------------------------------------------------------------
constexpr int MAXW = 68;

PVW_DEV int big_cmp(const u64* a, const u64* b, int n) {
  for (int i = n - 1; i >= 0; i--)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return 0;
}
PVW_DEV void big_sub(u64* r, const u64* a, const u64* b, int n) {  // r = a - b (mod 2^(64n))
  u64 borrow = 0;
  for (int i = 0; i < n; i++) {
    const u64 d1 = a[i] - b[i], b1 = a[i] < b[i];
    const u64 d2 = d1 - borrow, b2 = d1 < borrow;
    r[i] = d2;
    borrow = b1 | b2;
  }
}
PVW_DEV bool big_is_zero(const u64* a, int n) {
  for (int i = 0; i < n; i++)
    if (a[i]) return false;
  return true;
}
// (q, r) = (u1*2^64 + u0) / d, d normalised, u1 < d, v = floor((2^128-1)/d) - 2^64   (Moller-Granlund, Alg. 4)
PVW_DEV u64 div_2by1(u64 u1, u64 u0, u64 d, u64 v, u64* rem) {
  u64 q0 = v * u1, q1 = __umul64hi(v, u1);
  q0 += u0;
  q1 += u1 + (q0 < u0) + 1;
  u64 r = u0 - q1 * d;
  if (r > q0) { q1--; r += d; }
  if (r >= d) { q1++; r -= d; }
  *rem = r;
  return q1;
}
// Knuth algorithm D.  u: dividend, nu words (any value); v: normalised divisor (n words, top bit set), `shift` = the
// normalisation shift already applied to v.  q receives nq = nu_eff - n + 1 words (caller zero-fills q[0..qcap)),
// r receives the n-word remainder.  un is scratch of nu+1 words.
PVW_DEV void big_divrem(const u64* u, int nu, const u64* v, int n, int shift, u64 vinv, u64* q, u64* r, u64* un) {
  while (nu > 0 && u[nu - 1] == 0) nu--;
  if (nu < n) {  // |u| < |v|
    for (int i = 0; i < n; i++) r[i] = i < nu ? u[i] : 0;
    return;
  }
  un[nu] = shift ? u[nu - 1] >> (64 - shift) : 0;
  for (int i = nu - 1; i > 0; i--) un[i] = shift ? (u[i] << shift) | (u[i - 1] >> (64 - shift)) : u[i];
  un[0] = u[0] << shift;
  const u64 vt = v[n - 1];
  for (int j = nu - n; j >= 0; j--) {
    const u64 u1 = un[j + n], u0 = un[j + n - 1];
    u64 qhat, rhat;
    bool rhat_ovf = false;
    if (u1 >= vt) {  // only u1 == vt can occur; qhat = B - 1
      qhat = ~0ull;
      rhat = u0 + vt;
      rhat_ovf = rhat < u0;
    } else {
      qhat = div_2by1(u1, u0, vt, vinv, &rhat);
    }
    if (n >= 2) {
      const u64 v2 = v[n - 2], u2 = un[j + n - 2];
      while (!rhat_ovf) {
        const u64 ph = __umul64hi(qhat, v2), pl = qhat * v2;
        if (ph > rhat || (ph == rhat && pl > u2)) {
          qhat--;
          const u64 nr = rhat + vt;
          rhat_ovf = nr < rhat;
          rhat = nr;
        } else {
          break;
        }
      }
    }
    // multiply and subtract qhat * v from un[j .. j+n]
    u64 carry = 0, borrow = 0;
    for (int i = 0; i < n; i++) {
      const u64 pl = qhat * v[i], ph = __umul64hi(qhat, v[i]);
      const u64 lo = pl + carry;
      carry = ph + (lo < pl);
      const u64 x = un[i + j];
      const u64 d1 = x - lo, b1 = x < lo;
      const u64 d2 = d1 - borrow, b2 = d1 < borrow;
      un[i + j] = d2;
      borrow = b1 | b2;
    }
    {
      const u64 x = un[j + n];
      const u64 d1 = x - carry, b1 = x < carry;
      const u64 d2 = d1 - borrow, b2 = d1 < borrow;
      un[j + n] = d2;
      borrow = b1 | b2;
    }
    if (borrow) {  // add back
      qhat--;
      u64 c = 0;
      for (int i = 0; i < n; i++) {
        const u64 s1 = un[i + j] + v[i], c1 = s1 < v[i];
        const u64 s2 = s1 + c, c2 = s2 < c;
        un[i + j] = s2;
        c = c1 | c2;
      }
      un[j + n] += c;
    }
    q[j] = qhat;
  }
  for (int i = 0; i < n; i++) r[i] = shift ? (un[i] >> shift) | (un[i + 1] << (64 - shift)) : un[i];
}

__global__ void __launch_bounds__(128) decode_tail_kernel(const u64* __restrict__ X, uint64_t S, uint32_t Pc, uint32_t ell, u64* __restrict__ out,
                                                          size_t out_ps, const DevTables T, const FallbackList fb) {
  const uint64_t limit = fb.count ? *fb.count : S;
  for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < limit; it += (uint64_t)gridDim.x * blockDim.x) {
  const uint64_t s = fb.count ? fb.list[it] : it;
  const int NW = (int)T.NW;
  u64 noise[MAXW], cur[MAXW], num[MAXW], quo[MAXW], scratch[MAXW], rem[MAXW];
  // ---- last component: centred remainder modulo M = D^(l-1)   (reduce_modulo_poly, decryption.rs:154-178)
  for (int w = 0; w < NW; w++) cur[w] = X[((size_t)(ell - 1) * NW + w) * S + s];
  bool neg = big_cmp(cur, T.halfQ, NW) > 0;                      // centre(): x > Q/2  -> x - Q
  if (neg) big_sub(num, T.Qw, cur, NW); else for (int w = 0; w < NW; w++) num[w] = cur[w];
  for (int w = 0; w < NW + 2; w++) quo[w] = 0;
  for (int w = 0; w < NW; w++) rem[w] = 0;
  big_divrem(num, NW, T.divM_v, (int)T.divM_n, (int)T.divM_shift, T.divM_vinv, quo, rem, scratch);  // |x| % M (truncated)
  // `reduced > half_mod` / `reduced < -half_mod`: both mean |rem| > floor(M/2); the sign then flips
  bool rneg = neg;
  if (big_cmp(rem, T.halfM, NW) > 0) {
    big_sub(rem, T.Mw, rem, NW);  // M - |rem|
    rneg = !neg;
  }
  if (big_is_zero(rem, NW)) rneg = false;
  // noise_{l-1} as an element of [0,Q): bigints_to_poly of a signed value, parameters.rs:437-452
  if (rneg) big_sub(noise, T.Qw, rem, NW); else for (int w = 0; w < NW; w++) noise[w] = rem[w];
  // ---- back-substitution  noise_i = round((noise_{i+1} - tmp_i) / D)   (decryption.rs:44-48, :180-207)
  for (int i = (int)ell - 2; i >= 0; i--) {
    for (int w = 0; w < NW; w++) cur[w] = X[((size_t)i * NW + w) * S + s];
    // diff = (noise - tmp_i) mod Q
    if (big_cmp(noise, cur, NW) >= 0) big_sub(num, noise, cur, NW);
    else { big_sub(num, cur, noise, NW); big_sub(num, T.Qw, num, NW); }
    const bool nneg = big_cmp(num, T.halfQ, NW) > 0;
    if (nneg) big_sub(num, T.Qw, num, NW);                       // |centre(diff)|
    // |quotient| = floor((2|num| + D) / (2D)) for either sign (truncated division of 2num -/+ D by 2D)
    u64 c = 0;
    for (int w = 0; w < NW; w++) { const u64 x = num[w]; num[w] = (x << 1) | c; c = x >> 63; }
    num[NW] = c;
    c = 0;
    for (int w = 0; w <= NW; w++) {
      const u64 dw = w < NW ? T.Dw[w] : 0;
      const u64 s1 = num[w] + dw, c1 = s1 < dw;
      const u64 s2 = s1 + c, c2 = s2 < c;
      num[w] = s2;
      c = c1 | c2;
    }
    for (int w = 0; w < NW + 2; w++) quo[w] = 0;
    big_divrem(num, NW + 1, T.div2D_v, (int)T.div2D_n, (int)T.div2D_shift, T.div2D_vinv, quo, rem, scratch);
    if (nneg && !big_is_zero(quo, NW)) big_sub(noise, T.Qw, quo, NW); else for (int w = 0; w < NW; w++) noise[w] = quo[w];
  }
  // ---- plaintext = (-z_0) - noise_0 mod Q, centred, then extract_constant_term_as_u64 (decryption.rs:226-247)
  for (int w = 0; w < NW; w++) cur[w] = X[((size_t)ell * NW + w) * S + s];
  if (big_cmp(cur, noise, NW) >= 0) big_sub(num, cur, noise, NW);
  else { big_sub(num, noise, cur, NW); big_sub(num, T.Qw, num, NW); }
  u64 result;
  if (big_cmp(num, T.halfQ, NW) > 0) {           // negative: value = num - Q
    big_sub(cur, T.Qw, num, NW);                 // |value|
    bool small = cur[0] <= 1000;
    for (int w = 1; w < NW; w++) small = small && cur[w] == 0;
    if (small) result = 0;                       // "small negative values might be noise"
    else {                                       // (value + Q) % Q = num ; to_u64().unwrap_or(0)
      bool fits = true;
      for (int w = 1; w < NW; w++) fits = fits && num[w] == 0;
      result = fits ? num[0] : 0;
    }
  } else {
    bool fits = true;
    for (int w = 1; w < NW; w++) fits = fits && num[w] == 0;
    result = fits ? num[0] : 0;
  }
  const uint64_t d = s / Pc, p = s % Pc;
  out[p * out_ps + d] = result;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// (3') the same tail with every multi-precision value in registers: sizes are template parameters (NW words of Q,
//      NM words of M = D^(l-1), ND words of 2D), loops fully unrolled, constants broadcast from shared memory.
//      Quotient digits that are provably zero (top words of the running remainder) are skipped, so a share whose
//      noise is small costs a few digit steps while arbitrary inputs still take the full Knuth D path.
//      Bit-identical to decode_tail_kernel (tests/test_gpu_parity.py runs both on the same inputs).
// ---------------------------------------------------------------------------------------------------------------
struct TailConst {  // offsets (in u64 words) into the shared constant block
  int Q, halfQ, M, halfM, D, vM, v2D;
};

template <int N>
PVW_DEV bool gt_n(const u64 (&a)[N], const u64* b) {  // a > b
  bool gt = false, decided = false;
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    const u64 bw = b[i];
    if (!decided && a[i] != bw) { gt = a[i] > bw; decided = true; }
  }
  return gt;
}
template <int N>
PVW_DEV void rsub_n(u64 (&a)[N], const u64* b) {  // a = b - a
  u64 borrow = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const u64 bw = b[i], d1 = bw - a[i], b1 = bw < a[i], d2 = d1 - borrow, b2 = d1 < borrow;
    a[i] = d2;
    borrow = b1 | b2;
  }
}
template <int N>
PVW_DEV bool is_zero_n(const u64 (&a)[N]) {
  u64 o = 0;
#pragma unroll
  for (int i = 0; i < N; i++) o |= a[i];
  return o == 0;
}
// Knuth D on a normalised dividend un[NU+1] (in place; remainder left in un[0..NV), still shifted) by the normalised
// divisor v[NV]; q[NU-NV+1] receives the quotient.
template <int NU, int NV>
PVW_DEV void divrem_fixed(u64 (&un)[NU + 1], const u64* v, u64 vinv, u64 (&q)[NU - NV + 1]) {
  const u64 vt = v[NV - 1];
#pragma unroll
  for (int j = NU - NV; j >= 0; j--) {
    const u64 u1 = un[j + NV], u0 = un[j + NV - 1];
    if (u1 == 0 && u0 < vt) { q[j] = 0; continue; }  // digit is 0: (u1, u0, ...) < v
    u64 qhat, rhat;
    bool rhat_ovf = false;
    if (u1 >= vt) {
      qhat = ~0ull;
      rhat = u0 + vt;
      rhat_ovf = rhat < u0;
    } else {
      qhat = div_2by1(u1, u0, vt, vinv, &rhat);
    }
    if (NV >= 2) {
      const u64 v2 = v[NV - 2], u2 = un[j + NV - 2];
      while (!rhat_ovf) {
        const u64 ph = __umul64hi(qhat, v2), pl = qhat * v2;
        if (ph > rhat || (ph == rhat && pl > u2)) {
          qhat--;
          const u64 nr = rhat + vt;
          rhat_ovf = nr < rhat;
          rhat = nr;
        } else {
          break;
        }
      }
    }
    u64 carry = 0, borrow = 0;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const u64 vw = v[i];
      const u64 pl = qhat * vw, ph = __umul64hi(qhat, vw);
      const u64 lo = pl + carry;
      carry = ph + (lo < pl);
      const u64 x = un[i + j];
      const u64 d1 = x - lo, b1 = x < lo;
      const u64 d2 = d1 - borrow, b2 = d1 < borrow;
      un[i + j] = d2;
      borrow = b1 | b2;
    }
    {
      const u64 x = un[j + NV];
      const u64 d1 = x - carry, b1 = x < carry;
      const u64 d2 = d1 - borrow, b2 = d1 < borrow;
      un[j + NV] = d2;
      borrow = b1 | b2;
    }
    if (borrow) {
      qhat--;
      u64 c = 0;
#pragma unroll
      for (int i = 0; i < NV; i++) {
        const u64 vw = v[i];
        const u64 s1 = un[i + j] + vw, c1 = s1 < vw;
        const u64 s2 = s1 + c, c2 = s2 < c;
        un[i + j] = s2;
        c = c1 | c2;
      }
      un[j + NV] += c;
    }
    q[j] = qhat;
  }
}

template <int NW, int NM, int ND>
__global__ void __launch_bounds__(128, PVW_TAIL_MINB) decode_tail_fixed_kernel(const u64* __restrict__ X, uint64_t S, uint32_t Pc, uint32_t ell, u64* __restrict__ out,
                                                                size_t out_ps, const DevTables T, const FallbackList fb) {
  __shared__ u64 sc[5 * NW + NM + ND];
  const u64* cQ = sc; const u64* cHQ = sc + NW; const u64* cM = sc + 2 * NW; const u64* cHM = sc + 3 * NW; const u64* cD = sc + 4 * NW;
  const u64* cvM = sc + 5 * NW; const u64* cv2D = sc + 5 * NW + NM;
  for (int i = threadIdx.x; i < 5 * NW + NM + ND; i += blockDim.x) {
    const u64* src = i < NW ? T.Qw + i : i < 2 * NW ? T.halfQ + (i - NW) : i < 3 * NW ? T.Mw + (i - 2 * NW) : i < 4 * NW ? T.halfM + (i - 3 * NW)
                     : i < 5 * NW ? T.Dw + (i - 4 * NW) : i < 5 * NW + NM ? T.divM_v + (i - 5 * NW) : T.div2D_v + (i - 5 * NW - NM);
    sc[i] = *src;
  }
  __syncthreads();
  const int shM = (int)T.divM_shift, sh2D = (int)T.div2D_shift;
  const uint64_t limit = fb.count ? *fb.count : S;
  for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < limit; it += (uint64_t)gridDim.x * blockDim.x) {
  const uint64_t s = fb.count ? fb.list[it] : it;
  u64 noise[NW];
  // ---- last component: centred remainder modulo M (reduce_modulo_poly, decryption.rs:154-178)
  {
    u64 cur[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) cur[w] = X[((size_t)(ell - 1) * NW + w) * S + s];
    const bool neg = gt_n<NW>(cur, cHQ);
    if (neg) rsub_n<NW>(cur, cQ);                 // |centre(last)|
    u64 un[NW + 1];
    un[NW] = shM ? cur[NW - 1] >> (64 - shM) : 0;
#pragma unroll
    for (int i = NW - 1; i > 0; i--) un[i] = shM ? (cur[i] << shM) | (cur[i - 1] >> (64 - shM)) : cur[i];
    un[0] = cur[0] << shM;
    u64 q[NW - NM + 1];
    divrem_fixed<NW, NM>(un, cvM, T.divM_vinv, q);
    u64 rem[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) rem[i] = i < NM ? (shM ? (un[i] >> shM) | (un[i + 1] << (64 - shM)) : un[i]) : 0;
    bool rneg = neg;
    if (gt_n<NW>(rem, cHM)) { rsub_n<NW>(rem, cM); rneg = !neg; }
    if (is_zero_n<NW>(rem)) rneg = false;
    if (rneg) rsub_n<NW>(rem, cQ);                // signed value -> [0, Q) (bigints_to_poly, parameters.rs:437-452)
#pragma unroll
    for (int w = 0; w < NW; w++) noise[w] = rem[w];
  }
  // ---- back-substitution noise_i = round((noise_{i+1} - tmp_i) / D)  (decryption.rs:44-48, :180-207)
  for (int i = (int)ell - 2; i >= 0; i--) {
    u64 num[NW + 2];
    u64 borrow = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {                // noise - tmp_i
      const u64 c = X[((size_t)i * NW + w) * S + s];
      const u64 d1 = noise[w] - c, b1 = noise[w] < c, d2 = d1 - borrow, b2 = d1 < borrow;
      num[w] = d2;
      borrow = b1 | b2;
    }
    if (borrow) {                                  // ... mod Q
      u64 c = 0;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        const u64 qw = cQ[w], s1 = num[w] + qw, c1 = s1 < qw, s2 = s1 + c, c2 = s2 < c;
        num[w] = s2;
        c = c1 | c2;
      }
    }
    bool nneg;
    {
      u64 t[NW];
#pragma unroll
      for (int w = 0; w < NW; w++) t[w] = num[w];
      nneg = gt_n<NW>(t, cHQ);
      if (nneg) rsub_n<NW>(t, cQ);
      // 2|num| + D, NW+1 words, then normalise by the shift of 2D
      u64 c = 0, carry = 0;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        const u64 x = (t[w] << 1) | c;
        c = t[w] >> 63;
        const u64 dw = cD[w], s1 = x + dw, c1 = s1 < dw, s2 = s1 + carry, c2 = s2 < carry;
        num[w] = s2;
        carry = c1 | c2;
      }
      num[NW] = c + carry;
    }
    num[NW + 1] = sh2D ? num[NW] >> (64 - sh2D) : 0;
#pragma unroll
    for (int w = NW; w > 0; w--) num[w] = sh2D ? (num[w] << sh2D) | (num[w - 1] >> (64 - sh2D)) : num[w];
    num[0] = num[0] << sh2D;
    u64 q[NW + 1 - ND + 1];
    divrem_fixed<NW + 1, ND>(num, cv2D, T.div2D_vinv, q);
    // quotient < Q: NW words are enough
    u64 qq[NW];
    bool qzero = true;
#pragma unroll
    for (int w = 0; w < NW; w++) { qq[w] = w < NW + 1 - ND + 1 ? q[w] : 0; qzero = qzero && qq[w] == 0; }
    if (nneg && !qzero) rsub_n<NW>(qq, cQ);
#pragma unroll
    for (int w = 0; w < NW; w++) noise[w] = qq[w];
  }
  // ---- plaintext = (-z_0) - noise_0 mod Q, centred, then extract_constant_term_as_u64 (decryption.rs:226-247)
  u64 num[NW];
  {
    u64 borrow = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
      const u64 c = X[((size_t)ell * NW + w) * S + s];
      const u64 d1 = c - noise[w], b1 = c < noise[w], d2 = d1 - borrow, b2 = d1 < borrow;
      num[w] = d2;
      borrow = b1 | b2;
    }
    if (borrow) {
      u64 c = 0;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        const u64 qw = cQ[w], s1 = num[w] + qw, c1 = s1 < qw, s2 = s1 + c, c2 = s2 < c;
        num[w] = s2;
        c = c1 | c2;
      }
    }
  }
  u64 result;
  bool fits = true;
#pragma unroll
  for (int w = 1; w < NW; w++) fits = fits && num[w] == 0;
  if (gt_n<NW>(num, cHQ)) {                        // negative: value = num - Q
    u64 mag[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) mag[w] = num[w];
    rsub_n<NW>(mag, cQ);
    bool small = mag[0] <= 1000;
#pragma unroll
    for (int w = 1; w < NW; w++) small = small && mag[w] == 0;
    result = small ? 0 : (fits ? num[0] : 0);
  } else {
    result = fits ? num[0] : 0;
  }
  const uint64_t d = s / Pc, p = s % Pc;
  out[p * out_ps + d] = result;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// (0) FUSED FAST PATH -- one kernel for the shares a correct run produces, the three-kernel chain above / below only for the rest.
//
// For a share that decodes (z_i = -(m D^i + e_i) mod Q with every |e_i| < D/2; D = Delta) all the integers of the reference's
// procedure are small, and the procedure collapses (derivation in DESIGN.md 4 "Decode"):
//   t_i = centre(z_i D - z_{i+1}) = e_{i+1} - D e_i   fits a sub-basis of 3-4 limbs -> short lift, VERIFIED against all L residues;
//   H   = sum_i t_i D^(l-2-i) is `last` as an integer provided |H| <= Q/2.  Writing  t_{l-2} = d_0 + D x_0  (|d_0| <= D/2)  and
//         t_{l-2-j} + x_{j-1} = D x_j  EXACTLY for j = 1..l-2  gives  H = d_0 + x_{l-2} D^(l-1):  the centred remainder of `last`
//         modulo D^(l-1) is d_0 and every rounded division of the back-substitution is exact: noise_i = -x_{l-2-i};
//   m   = centre(-z_0) + x_{l-2}   with centre(-z_0) short-lifted and verified like the t_i.
// Every claim is CHECKED per share (verification of each lift against all residues, exact divisibility, |x_{l-2}| <= cmax so that
// |H| <= Q/2, the u64 rules); a share that fails any check is appended to a list and takes the general chain, which computes the
// reference's result for arbitrary inputs.  So results are bit-identical for every input, and a correct run moves z and c2
// through HBM once: no y / X round trips, no full CRT lift, no long division.
//   phase A: thread = (share, limb): inverse NTT, t_i and -z_0 residues -> shared memory
//   phase B: thread = (share, value): short lift + verification -> shared memory
//   phase C: thread = share: the l-1 exact divisions by D on numbers of at most four words
// ---------------------------------------------------------------------------------------------------------------
struct Small { u64 m[4]; bool neg; };   // sign + magnitude, |value| < 2^255

// (q, r) = |c| / D for a normalised divisor of ND words, when the quotient fits one word; false otherwise
template <int ND>
PVW_DEV bool div_small(const u64 (&mag)[4], const FusedConst& F, u64& q, bool& rem_zero, bool& rem_gt_half, u64* rem_out = nullptr) {
  const int sh = (int)F.shift;
  u64 un[5];
  un[4] = sh ? mag[3] >> (64 - sh) : 0;
#pragma unroll
  for (int i = 3; i > 0; i--) un[i] = sh ? (mag[i] << sh) | (mag[i - 1] >> (64 - sh)) : mag[i];
  un[0] = mag[0] << sh;
  bool ok = true;
#pragma unroll
  for (int w = ND + 1; w < 5; w++) ok = ok && un[w] == 0;
  const u64 vt = F.dv[ND - 1];
  if (!ok || un[ND] >= vt) return false;                    // quotient of two or more words
  u64 rhat, qhat = div_2by1(un[ND], un[ND - 1], vt, F.vinv, &rhat);
  if (ND >= 2) {
    const u64 v2 = F.dv[ND - 2], u2 = un[ND - 2];
    bool ovf = false;
    while (!ovf) {
      const u64 ph = __umul64hi(qhat, v2), pl = qhat * v2;
      if (ph > rhat || (ph == rhat && pl > u2)) {
        qhat--;
        const u64 nr = rhat + vt;
        ovf = nr < rhat;
        rhat = nr;
      } else {
        break;
      }
    }
  }
  u64 carry = 0, borrow = 0;
#pragma unroll
  for (int i = 0; i < ND; i++) {
    const u64 vw = F.dv[i], pl = qhat * vw, ph = __umul64hi(qhat, vw), lo = pl + carry;
    carry = ph + (lo < pl);
    const u64 x = un[i], d1 = x - lo, b1 = x < lo, d2 = d1 - borrow, b2 = d1 < borrow;
    un[i] = d2;
    borrow = b1 | b2;
  }
  {
    const u64 x = un[ND], d1 = x - carry, b1 = x < carry, d2 = d1 - borrow, b2 = d1 < borrow;
    un[ND] = d2;
    borrow = b1 | b2;
  }
  if (borrow) {
    qhat--;
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < ND; i++) {
      const u64 vw = F.dv[i], s1 = un[i] + vw, c1 = s1 < vw, s2 = s1 + c, c2 = s2 < c;
      un[i] = s2;
      c = c1 | c2;
    }
    un[ND] += c;
  }
  q = qhat;
  u64 r[ND];
  u64 any = 0;
#pragma unroll
  for (int i = 0; i < ND; i++) { r[i] = sh ? (un[i] >> sh) | (un[i + 1] << (64 - sh)) : un[i]; any |= r[i]; }
  if (rem_out) {
#pragma unroll
    for (int i = 0; i < 4; i++) rem_out[i] = 0;
#pragma unroll
    for (int i = 0; i < ND; i++) rem_out[i] = r[i];
  }
  rem_zero = any == 0;
  bool gt = false, decided = false;
#pragma unroll
  for (int i = ND - 1; i >= 0; i--)
    if (!decided && r[i] != F.half_d[i]) { gt = r[i] > F.half_d[i]; decided = true; }
  rem_gt_half = gt;
  return true;
}

// a + sign * x for a 4-word signed a and a one-word magnitude x
PVW_DEV Small add_small(const Small& a, u64 x, bool xneg) {
  Small r;
  if (a.neg == xneg) {
    u64 c = x;
#pragma unroll
    for (int i = 0; i < 4; i++) { const u64 s1 = a.m[i] + c; c = s1 < c; r.m[i] = s1; }
    r.neg = a.neg;
  } else {
    const bool a_ge = (a.m[3] | a.m[2] | a.m[1]) != 0 || a.m[0] >= x;
    if (a_ge) {
      u64 b = x;
#pragma unroll
      for (int i = 0; i < 4; i++) { const u64 d1 = a.m[i] - b; b = a.m[i] < b; r.m[i] = d1; }
      r.neg = a.neg;
    } else {
      r.m[0] = x - a.m[0]; r.m[1] = r.m[2] = r.m[3] = 0;
      r.neg = xneg;
    }
  }
  if ((r.m[0] | r.m[1] | r.m[2] | r.m[3]) == 0) r.neg = false;
  return r;
}

template <int ELL, int G, int SW, int ND>
__global__ void __launch_bounds__(128) decode_fused_kernel(const u64* __restrict__ z, size_t z_ls, size_t z_ds, size_t z_cs, const DecodeSub sub,
                                                           uint32_t Pc, uint64_t S, u64* __restrict__ out, size_t out_ps, const DevTables T,
                                                           const FusedConst F, uint32_t* __restrict__ fb_list, uint32_t* __restrict__ fb_count) {
  extern __shared__ __align__(16) u64 sm[];
  const uint32_t L = T.L;
  u64* s_twi = sm;                                   // [L][ELL]
  u64* s_twi_sh = s_twi + (size_t)L * ELL;           // [L][ELL]
  u64* s_vc = s_twi_sh + (size_t)L * ELL;            // [L][10]  q, floor(2^64/q), -, -, 2^64 / 2^128 / 2^192 mod q + Shoup
  u64* s_dc = s_vc + (size_t)L * 10;                 // [L][4]   decode_rns multipliers
  u64* Y = s_dc + (size_t)L * 4;                     // [ELL][L][G]  residues of t_0..t_{l-2}, -z_0
  u64* Tm = Y + (size_t)ELL * L * G;                 // [ELL][4][G]  lifted magnitudes
  uint32_t* Tf = reinterpret_cast<uint32_t*>(Tm + (size_t)ELL * 4 * G);   // [ELL][G]  bit 0: negative, bit 1: lift verified
  for (uint32_t i = threadIdx.x; i < L * ELL; i += blockDim.x) { s_twi[i] = T.twi[i]; s_twi_sh[i] = T.twi_sh[i]; }
  for (uint32_t j = threadIdx.x; j < L; j += blockDim.x) {
    u64* cj = s_vc + (size_t)j * 10;
    cj[0] = T.lc[j].q; cj[1] = T.lc[j].mu64; cj[2] = 0; cj[3] = 0;
    for (int t = 0; t < 3; t++) { cj[4 + 2 * t] = T.sh_r[(size_t)j * 3 + t]; cj[5 + 2 * t] = T.sh_r_sh[(size_t)j * 3 + t]; }
    for (int t = 0; t < 4; t++) s_dc[(size_t)j * 4 + t] = T.dec_c[(size_t)j * 4 + t];
  }
  __syncthreads();
  const uint64_t ngroups = (S + G - 1) / G;
  for (uint64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const uint64_t s0 = grp * G;
    // ---- phase A
    for (uint32_t item = threadIdx.x; item < L * G; item += blockDim.x) {
      const uint32_t limb = item / G, g = item - limb * G;
      const uint64_t s = s0 + g;
      if (s >= S) continue;
      const uint64_t d = s / Pc, p = s % Pc;
      const u64 q = s_vc[(size_t)limb * 10];
      u64 a[ELL];
      if (z_cs) {
        const u64* src = z + d * z_ds + (size_t)limb * z_ls + p;
#pragma unroll
        for (int t = 0; t < ELL; t++) a[t] = src[(size_t)t * z_cs];
      } else {
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(z + d * z_ds + (size_t)limb * z_ls + p * ELL);
#pragma unroll
        for (int t = 0; t < ELL / 2; t++) { const ulonglong2 v = src[t]; a[2 * t] = v.x; a[2 * t + 1] = v.y; }
      }
      if (sub.S) {
        const uint32_t sd = sub.dmap ? sub.dmap[d] : (uint32_t)d, srow = sub.rowmap ? sub.rowmap[p] : (uint32_t)p;
        const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(sub.S + (size_t)sd * sub.S_ds + (size_t)limb * sub.S_ls + (size_t)srow * ELL);
#pragma unroll
        for (int t = 0; t < ELL / 2; t++) {
          const ulonglong2 v = sp[t];
          a[2 * t] = submod(a[2 * t], v.x, q);
          a[2 * t + 1] = submod(a[2 * t + 1], v.y, q);
        }
      }
      ntt_inverse_unscaled_regs<ELL>(a, s_twi + (size_t)limb * ELL, s_twi_sh + (size_t)limb * ELL, q);
      const u64 c1 = s_dc[4 * limb], c1_sh = s_dc[4 * limb + 1], c2 = s_dc[4 * limb + 2], c2_sh = s_dc[4 * limb + 3];
      u64* yo = Y + (size_t)limb * G + g;
      u64 p2 = mulmod_shoup(a[0], c2, c2_sh, q);
      yo[(size_t)(ELL - 1) * L * G] = negmod(p2, q);                                    // -z_0
#pragma unroll
      for (int i = 0; i < ELL - 1; i++) {
        p2 = mulmod_shoup(a[i + 1], c2, c2_sh, q);
        yo[(size_t)i * L * G] = submod(mulmod_shoup(a[i], c1, c1_sh, q), p2, q);        // z_i D - z_{i+1}
      }
    }
    __syncthreads();
    // ---- phase B
    for (uint32_t item = threadIdx.x; item < ELL * G; item += blockDim.x) {
      const uint32_t i = item / G, g = item - i * G;
      if (s0 + g >= S) continue;
      u64 mag[4];
      bool neg = false;
      const bool ok = short_lift<SW>(Y + (size_t)i * L * G + g, G, L, T, s_vc, mag, neg);
#pragma unroll
      for (int w = 0; w < 4; w++) Tm[((size_t)i * 4 + w) * G + g] = mag[w];
      Tf[(size_t)i * G + g] = (neg ? 1u : 0u) | (ok ? 2u : 0u);
    }
    __syncthreads();
    // ---- phase C
    if (threadIdx.x < G && s0 + threadIdx.x < S) {
      const uint32_t g = threadIdx.x;
      const uint64_t s = s0 + g;
      auto value = [&](int i) {
        Small v;
#pragma unroll
        for (int w = 0; w < 4; w++) v.m[w] = Tm[((size_t)i * 4 + w) * G + g];
        v.neg = (Tf[(size_t)i * G + g] & 1u) != 0;
        return v;
      };
      bool ok = true;
#pragma unroll 1
      for (int i = 0; i < ELL; i++) ok = ok && (Tf[(size_t)i * G + g] & 2u) != 0;
      u64 x = 0;
      bool xneg = false;
      if (ok) {
        // t_{l-2} = d_0 + D x_0 with the centred remainder d_0
        const Small t = value(ELL - 2);
        u64 q0;
        bool rz, rgh;
        ok = div_small<ND>(t.m, F, q0, rz, rgh);
        if (ok && rgh) { q0++; ok = q0 != 0; }
        x = q0; xneg = t.neg && q0 != 0;
#pragma unroll 1
        for (int j = 1; j <= ELL - 2 && ok; j++) {
          const Small c = add_small(value(ELL - 2 - j), x, xneg);   // t_{l-2-j} + x_{j-1} must be an exact multiple of D
          u64 qj;
          ok = div_small<ND>(c.m, F, qj, rz, rgh) && rz;
          x = qj; xneg = c.neg && qj != 0;
        }
        ok = ok && x <= F.cmax;                                      // |H| = |d_0 + x D^(l-1)| <= Q/2: `last` did not wrap
      }
      u64 result = 0;
      if (ok) {
        const Small pt = add_small(value(ELL - 1), x, xneg);         // plaintext = -z_0 - noise_0 = centre(-z_0) + x_{l-2}
        const bool hi = (pt.m[1] | pt.m[2] | pt.m[3]) != 0;
        if (pt.neg) {
          if (!hi && pt.m[0] <= 1000) result = 0;                    // "small negative values might be noise", decryption.rs:226-247
          else ok = false;                                           // (pt + Q) % Q then to_u64: left to the general path
        } else {
          result = hi ? 0 : pt.m[0];                                 // to_u64().unwrap_or(0)
        }
      }
      if (ok) {
        const uint64_t d = s / Pc, p = s % Pc;
        out[p * out_ps + d] = result;
      } else {
        fb_list[atomicAdd(fb_count, 1u)] = (uint32_t)s;
      }
    }
    __syncthreads();
  }
}

// Residues of t_0..t_{l-2} and -z_0 of one share in one limb (the per-limb step shared by the kernels below).
template <int ELL, bool LAZY = false>   // LAZY: results only below 4q (callers that multiply them anyway), Harvey butterflies
PVW_DEV void share_residues(const u64* __restrict__ z, size_t z_ls, size_t z_ds, size_t z_cs, const DecodeSub& sub, uint64_t d, uint64_t p, uint32_t sd,
                            uint32_t srow, uint32_t limb, u64 q, const u64* tw, const u64* tw_sh, const u64* dc, u64 (&y)[ELL]) {
  u64 a[ELL];
  if (z_cs) {
    const u64* src = z + d * z_ds + (size_t)limb * z_ls + p;
#pragma unroll
    for (int t = 0; t < ELL; t++) a[t] = src[(size_t)t * z_cs];
  } else {
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(z + d * z_ds + (size_t)limb * z_ls + p * ELL);
#pragma unroll
    for (int t = 0; t < ELL / 2; t++) { const ulonglong2 v = src[t]; a[2 * t] = v.x; a[2 * t + 1] = v.y; }
  }
  if (sub.S) {
    const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(sub.S + (size_t)sd * sub.S_ds + (size_t)limb * sub.S_ls + (size_t)srow * ELL);
#pragma unroll
    for (int t = 0; t < ELL / 2; t++) {
      const ulonglong2 v = sp[t];
      a[2 * t] = LAZY ? a[2 * t] - v.x + q : submod(a[2 * t], v.x, q);
      a[2 * t + 1] = LAZY ? a[2 * t + 1] - v.y + q : submod(a[2 * t + 1], v.y, q);
    }
  }
  const u64 c1 = dc[0], c1_sh = dc[1], c2 = dc[2], c2_sh = dc[3];
  if (LAZY) {
    ntt_inverse_unscaled_lazy_regs<ELL>(a, tw, tw_sh, q);
    const u64 q2 = 2 * q;
    u64 p2 = mulmod_shoup_lazy(a[0], c2, c2_sh, q);
    y[ELL - 1] = q2 - p2;                                                       // -z_0, in (0, 2q]
#pragma unroll
    for (int i = 0; i < ELL - 1; i++) {
      p2 = mulmod_shoup_lazy(a[i + 1], c2, c2_sh, q);
      y[i] = mulmod_shoup_lazy(a[i], c1, c1_sh, q) - p2 + q2;                   // z_i D - z_{i+1}, in (0, 4q)
    }
    return;
  }
  ntt_inverse_unscaled_regs<ELL>(a, tw, tw_sh, q);
  u64 p2 = mulmod_shoup(a[0], c2, c2_sh, q);
  y[ELL - 1] = negmod(p2, q);                                                   // -z_0
#pragma unroll
  for (int i = 0; i < ELL - 1; i++) {
    p2 = mulmod_shoup(a[i + 1], c2, c2_sh, q);
    y[i] = submod(mulmod_shoup(a[i], c1, c1_sh, q), p2, q);                     // z_i D - z_{i+1}
  }
}

// The default form (l = 8, 16): thread = share, a loop over the limbs, nothing but constants in shared memory, no barrier, every
// global load unit stride across the warp.  (The staged kernel above ran at 2.5 ms per bench step -- phases serialised behind
// barriers, one warp in four busy during phase C -- and a thread-per-share form that verified every lifted value in every limb at
// 2.25 ms; this one: 1.4 ms.  The staged kernel is kept for l = 32, whose l * (SW + 1) accumulator words do not fit the registers.)
// Pass 1 walks the sub-basis limbs and accumulates the short lift of the l values; the carry chain then runs BEFORE the other
// limbs are looked at and yields the share's message m and noise e_0..e_{l-1} (one-word integers); pass 2 checks the CLAIM
//     z_i = -(m D^i + e_i)  (mod q_j)   for every remaining limb j and every i
// on the unscaled inverse transform directly: one Shoup multiply (m * (l D^i)) and one word reduction (l e_i) per coefficient
// instead of forming t_i, -z_0 in that limb and reducing a multi-word candidate (51 -> 29 modular multiplies per limb).
// Why the claim suffices: modulo the sub-basis it holds by construction (t_i = e_{i+1} - D e_i and centre(-z_0) = m + e_0 as integers,
// by induction on i); with the check it holds modulo Q, so tmp_i = e_{i+1} - D e_i are the centred values the reference works
// with (|.| < Q/2), H = e_{l-1} - D^(l-1) e_0 telescopes, |e_{l-1}| <= D/2 (centred remainder) gives red = e_{l-1}, every rounded
// division of the back-substitution is exact, noise_0 = e_0 and the plaintext is centre(-z_0 - e_0) = m.
// PHASE 1 / 2: the procedure runs as two launches -- pass 1 and the carry chain (the
// register-hungry half: l * (SW + 1) accumulator words) leave |e_i|, |m| and the signs in a scratch of l + 2 words per share, and the
// claim check of the remaining limbs (72 % of the instructions) runs as its own kernel at <= 80 registers, 24 instead of 16 warps
// per SM, from 1 000 instead of 9 700 instructions.  (PHASE 0, everything in one kernel, is the form this grew out of: 1.39 against
// 1.18 ms per million shares, a fifth of its cycles waiting for instruction fetch; it is no longer instantiated -- a third of this
// file's compile time -- but the template still describes it.)
template <int ELL, int SW, int ND, int MINB, int PHASE>
__global__ void __launch_bounds__(128, MINB) decode_fused_claim_kernel(const u64* __restrict__ z, size_t z_ls, size_t z_ds, size_t z_cs, const DecodeSub sub,
                                                                       uint32_t Pc, uint64_t S, u64* __restrict__ out, size_t out_ps, const DevTables T,
                                                                       const FusedConst F, uint32_t* __restrict__ fb_list, uint32_t* __restrict__ fb_count, u64* __restrict__ scr) {
  extern __shared__ __align__(16) u64 sm[];
  const uint32_t L = T.L, Ls = T.shortL;
  u64* s_twi = sm;                                   // [L][ELL]
  u64* s_twi_sh = s_twi + (size_t)L * ELL;           // [L][ELL]
  u64* s_lg = s_twi_sh + (size_t)L * ELL;            // [L][ELL] l D^i mod q
  u64* s_lg_sh = s_lg + (size_t)L * ELL;             // [L][ELL]
  u64* s_vc = s_lg_sh + (size_t)L * ELL;             // [L][4]   q, floor(2^64/q), sh_c, sh_c_sh (sub-basis limbs)
  u64* s_dc = s_vc + (size_t)L * 4;                  // [L][4]   decode_rns multipliers
  u64* s_qh = s_dc + (size_t)L * 4;                  // [Ls][SW] Q_s / q_j, then Q_s [SW], floor(Q_s / 2) [SW]
  for (uint32_t i = threadIdx.x; i < L * ELL; i += blockDim.x) { s_twi[i] = T.twi[i]; s_twi_sh[i] = T.twi_sh[i]; s_lg[i] = T.lgad[i]; s_lg_sh[i] = T.lgad_sh[i]; }
  for (uint32_t j = threadIdx.x; j < L; j += blockDim.x) {
    u64* cj = s_vc + (size_t)j * 4;
    cj[0] = T.lc[j].q; cj[1] = T.lc[j].mu64;
    cj[2] = j < Ls ? T.sh_c[j] : 0; cj[3] = j < Ls ? T.sh_c_sh[j] : 0;
    for (int t = 0; t < 4; t++) s_dc[(size_t)j * 4 + t] = T.dec_c[(size_t)j * 4 + t];
  }
  for (uint32_t i = threadIdx.x; i < Ls * SW; i += blockDim.x) s_qh[i] = T.sh_qhat[i];
  for (uint32_t i = threadIdx.x; i < SW; i += blockDim.x) { s_qh[Ls * SW + i] = T.sh_Q[i]; s_qh[Ls * SW + SW + i] = T.sh_halfQ[i]; }
  __syncthreads();
  const u64* sQ = s_qh + Ls * SW;
  const u64* sHQ = sQ + SW;
  const u64 E_MAX = F.emax;                            // |e_i| up to this: l * e_i < q_j for every limb, no reduction needed
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t d = s / Pc, p = s % Pc;
    const uint32_t sd = sub.S ? (sub.dmap ? sub.dmap[d] : (uint32_t)d) : 0, srow = sub.S ? (sub.rowmap ? sub.rowmap[p] : (uint32_t)p) : 0;
    u64 em[ELL];          // |e_i|
    uint32_t eneg = 0;    // sign bits of e_i
    bool ok = true;
    u64 mmag = 0, result = 0;
    bool mneg = false;
    if constexpr (PHASE != 2) {
    // ---- pass 1: short lift of t_0..t_{l-2}, -z_0 over the sub-basis
    u64 acc[ELL][SW + 1];
#pragma unroll
    for (int i = 0; i < ELL; i++)
#pragma unroll
      for (int w = 0; w <= SW; w++) acc[i][w] = 0;
#pragma unroll 1
    for (uint32_t j = 0; j < Ls; j++) {
      const u64* cj = s_vc + (size_t)j * 4;
      const u64 q = cj[0];
      u64 y[ELL];
      share_residues<ELL, true>(z, z_ls, z_ds, z_cs, sub, d, p, sd, srow, j, q, s_twi + (size_t)j * ELL, s_twi_sh + (size_t)j * ELL, s_dc + 4 * j, y);
      u64 qw[SW];
#pragma unroll
      for (int w = 0; w < SW; w++) qw[w] = s_qh[j * SW + w];
#pragma unroll
      for (int i = 0; i < ELL; i++) {
        const u64 t = mulmod_shoup(y[i], cj[2], cj[3], q);
        u64 carry = 0;
#pragma unroll
        for (int w = 0; w < SW; w++) {
          const u64 lo = t * qw[w], hi = __umul64hi(t, qw[w]);
          u64 x = acc[i][w] + carry;
          const u64 k1 = x < carry;
          x += lo;
          const u64 k2 = x < lo;
          acc[i][w] = x;
          carry = hi + k1 + k2;
        }
        acc[i][SW] += carry;
      }
    }
    uint32_t negs = 0;
#pragma unroll
    for (int i = 0; i < ELL; i++) {
      for (uint32_t r = 1; r < Ls; r++) {
        bool ge = acc[i][SW] != 0;
        if (!ge) {
          ge = true;
#pragma unroll
          for (int w = SW - 1; w >= 0; w--)
            if (acc[i][w] != sQ[w]) { ge = acc[i][w] > sQ[w]; break; }
        }
        if (ge) {
          u64 borrow = 0;
#pragma unroll
          for (int w = 0; w < SW; w++) {
            const u64 qv = sQ[w], d1 = acc[i][w] - qv, b1 = acc[i][w] < qv, d2 = d1 - borrow, b2 = d1 < borrow;
            acc[i][w] = d2;
            borrow = b1 | b2;
          }
          acc[i][SW] -= borrow;
        }
      }
      bool neg = false;
#pragma unroll
      for (int w = SW - 1; w >= 0; w--)
        if (acc[i][w] != sHQ[w]) { neg = acc[i][w] > sHQ[w]; break; }
      if (neg) {
        u64 borrow = 0;
#pragma unroll
        for (int w = 0; w < SW; w++) {
          const u64 qv = sQ[w], d1 = qv - acc[i][w], b1 = qv < acc[i][w], d2 = d1 - borrow, b2 = d1 < borrow;
          acc[i][w] = d2;
          borrow = b1 | b2;
        }
        negs |= 1u << i;
      }
    }
    // ---- the carry chain on the candidates: e_{l-1} = d_0, e_{l-2-j} = -x_j, m = centre(-z_0) + x_{l-2}
    u64 x = 0;
    bool xneg = false;
#pragma unroll
    for (int j = 0; j <= ELL - 2; j++) {
      Small c;
#pragma unroll
      for (int w = 0; w < 4; w++) c.m[w] = w < SW ? acc[ELL - 2 - j][w] : 0;
      c.neg = ((negs >> (ELL - 2 - j)) & 1u) != 0;
      if (j > 0) c = add_small(c, x, xneg);
      u64 qj, rem[4];
      bool rz, rgh;
      ok = div_small<ND>(c.m, F, qj, rz, rgh, rem) && ok;
      if (j == 0) {
        // centred remainder d_0 = e_{l-1}: r, or r - D with the quotient one up
        bool dneg = c.neg;
        if (rgh) {
          qj++;
          ok = ok && qj != 0;
          u64 borrow = 0;
#pragma unroll
          for (int w = 0; w < 4; w++) {           // D - r  (F.half_d holds floor(D/2); D itself = dv >> shift)
            const u64 dw = w < ND ? (F.shift ? (F.dv[w] >> F.shift) | ((w + 1 < ND ? F.dv[w + 1] : 0ull) << (64 - F.shift)) : F.dv[w]) : 0ull;
            const u64 d1 = dw - rem[w], b1 = dw < rem[w], d2 = d1 - borrow, b2 = d1 < borrow;
            rem[w] = d2;
            borrow = b1 | b2;
          }
          dneg = !dneg;
        }
        ok = ok && (rem[1] | rem[2] | rem[3]) == 0 && rem[0] <= E_MAX;
        em[ELL - 1] = rem[0];
        if (dneg && rem[0] != 0) eneg |= 1u << (ELL - 1);
      } else {
        ok = ok && rz;
      }
      x = qj; xneg = c.neg && qj != 0;
      ok = ok && x <= E_MAX;
      em[ELL - 2 - j] = x;
      if (!xneg && x != 0) eneg |= 1u << (ELL - 2 - j);      // e = -x
    }
    ok = ok && x <= F.cmax;                                   // |e_0| D^(l-1) + |e_{l-1}| <= Q/2: `last` does not wrap
    {
      Small w;
#pragma unroll
      for (int k = 0; k < 4; k++) w.m[k] = k < SW ? acc[ELL - 1][k] : 0;
      w.neg = ((negs >> (ELL - 1)) & 1u) != 0;
      const Small pt = add_small(w, x, xneg);
      ok = ok && (pt.m[1] | pt.m[2] | pt.m[3]) == 0;          // a plaintext of two or more words: to_u64 fails -> general path
      mmag = pt.m[0]; mneg = pt.neg;
      if (pt.neg) ok = ok && pt.m[0] <= 1000;                 // small negative -> 0 (decryption.rs:226-247); larger ones: general path
      result = pt.neg ? 0 : pt.m[0];
    }
    }  // PHASE != 2
    if constexpr (PHASE == 1) {
      // scratch, word-major (unit stride across the threads): |e_0| .. |e_{l-1}|, |m|, signs + verdict
#pragma unroll
      for (int i = 0; i < ELL; i++) scr[(size_t)i * S + s] = em[i];
      scr[(size_t)ELL * S + s] = mmag;
      scr[(size_t)(ELL + 1) * S + s] = (u64)eneg | ((u64)(mneg ? 1 : 0) << 32) | ((u64)(ok ? 1 : 0) << 33);
      if (!ok) fb_list[atomicAdd(fb_count, 1u)] = (uint32_t)s;
      continue;
    }
    if constexpr (PHASE == 2) {
      const u64 meta = scr[(size_t)(ELL + 1) * S + s];
      if (!((meta >> 33) & 1)) continue;                    // listed by the first launch already
      eneg = (uint32_t)meta; mneg = ((meta >> 32) & 1) != 0;
      mmag = scr[(size_t)ELL * S + s];
      result = mneg ? 0 : mmag;
#pragma unroll
      for (int i = 0; i < ELL; i++) em[i] = scr[(size_t)i * S + s];
    }
    // ---- pass 2: the claim in every other limb
#pragma unroll 1
    for (uint32_t j = Ls; j < L && ok; j++) {
      const u64* cj = s_vc + (size_t)j * 4;
      const u64 q = cj[0], mu = cj[1];
      u64 a[ELL];
      if (z_cs) {
        const u64* src = z + d * z_ds + (size_t)j * z_ls + p;
#pragma unroll
        for (int t = 0; t < ELL; t++) a[t] = src[(size_t)t * z_cs];
      } else {
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(z + d * z_ds + (size_t)j * z_ls + p * ELL);
#pragma unroll
        for (int t = 0; t < ELL / 2; t++) { const ulonglong2 v = src[t]; a[2 * t] = v.x; a[2 * t + 1] = v.y; }
      }
      if (sub.S) {
        const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(sub.S + (size_t)sd * sub.S_ds + (size_t)j * sub.S_ls + (size_t)srow * ELL);
#pragma unroll
        for (int t = 0; t < ELL / 2; t++) {
          const ulonglong2 v = sp[t];
          a[2 * t] = a[2 * t] - v.x + q;                        // (0, 2q): the lazy transform takes it
          a[2 * t + 1] = a[2 * t + 1] - v.y + q;
        }
      }
      ntt_inverse_unscaled_lazy_regs<ELL>(a, s_twi + (size_t)j * ELL, s_twi_sh + (size_t)j * ELL, q);   // l z_i, in [0, 2q)
      u64 mr = mmag - __umul64hi(mmag, mu) * q;                  // |m| mod q
      mr = csub(csub(mr, 2 * q), q);
      const u64* lg = s_lg + (size_t)j * ELL;
      const u64* lg_sh = s_lg_sh + (size_t)j * ELL;
      bool good = true;
#pragma unroll
      for (int i = 0; i < ELL; i++) {
        u64 t1 = mulmod_shoup(mr, lg[i], lg_sh[i], q);           // |m| l D^i
        if (mneg) t1 = t1 ? q - t1 : 0;
        u64 t2 = em[i] * (u64)ELL;                               // l |e_i| < q (emax)
        if ((eneg >> i) & 1u) t2 = t2 ? q - t2 : 0;
        const u64 sgm = csub(t1 + t2, q);                        // l (m D^i + e_i) mod q
        const u64 ai = csub(a[i], q);
        good = good && ((ai + sgm == q) || ((ai | sgm) == 0));   // l z_i = -(...)
      }
      ok = good;
    }
    if (ok) out[p * out_ps + d] = result;
    else fb_list[atomicAdd(fb_count, 1u)] = (uint32_t)s;
  }
}

template <int ELL, int SW, int MINB>
static bool launch_fused_claim(const DevTables& T, const FusedConst& F, const u64* z, size_t z_ls, size_t z_ds, size_t z_cs, const DecodeSub& sb, uint32_t Pc,
                               uint64_t S, u64* out, size_t out_ps, uint32_t* fb_list, uint32_t* fb_count, cudaStream_t st, u64* scr) {
  const size_t smem = ((size_t)T.L * ELL * 4 + (size_t)T.L * 8 + (size_t)T.shortL * SW + 2 * SW) * 8;
  if (smem > 96 * 1024 || scr == nullptr) return false;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
  // one launch of a phase: persistent-style grid, four CTAs' worth of shares per resident CTA
  auto run = [&](auto kern) {
    int per_sm = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem) != cudaSuccess || per_sm < 1) return false;
    const unsigned grid = (unsigned)std::min<uint64_t>((S + 127) / 128, (uint64_t)sms * per_sm * 4);
    kern<<<grid, 128, smem, st>>>(z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, T, F, fb_list, fb_count, scr);
    return true;
  };
#define PVW_FUSED_CLAIM_ND(N)                                                                                                    \
  case N:                                                                                                                        \
    return run(decode_fused_claim_kernel<ELL, SW, N, MINB, 1>) && run(decode_fused_claim_kernel<ELL, SW, N, (ELL == 8 ? PVW_FUSED_P2_MINB : 4), 2>);
  switch (F.nd) {
    PVW_FUSED_CLAIM_ND(1)
    PVW_FUSED_CLAIM_ND(2)
    PVW_FUSED_CLAIM_ND(3)
    PVW_FUSED_CLAIM_ND(4)
  }
#undef PVW_FUSED_CLAIM_ND
  return false;
}

template <int ELL, int MINB>
static bool launch_fused_claim_sw(const DevTables& T, const FusedConst& F, const u64* z, size_t z_ls, size_t z_ds, size_t z_cs, const DecodeSub& sb, uint32_t Pc,
                                  uint64_t S, u64* out, size_t out_ps, uint32_t* fb_list, uint32_t* fb_count, cudaStream_t st, u64* scr) {
  switch (T.shortSW) {
    case 2: return launch_fused_claim<ELL, 2, MINB>(T, F, z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, fb_list, fb_count, st, scr);
    case 3: return launch_fused_claim<ELL, 3, MINB>(T, F, z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, fb_list, fb_count, st, scr);
    case 4: return launch_fused_claim<ELL, 4, MINB>(T, F, z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, fb_list, fb_count, st, scr);
  }
  return false;
}

template <int ELL, int G, int SW>
static bool launch_fused_nd(const DevTables& T, const FusedConst& F, const u64* z, size_t z_ls, size_t z_ds, size_t z_cs, const DecodeSub& sb, uint32_t Pc,
                            uint64_t S, u64* out, size_t out_ps, uint32_t* fb_list, uint32_t* fb_count, cudaStream_t st) {
  const size_t smem = ((size_t)T.L * ELL * 2 + (size_t)T.L * 14 + (size_t)ELL * T.L * G + (size_t)ELL * 4 * G) * 8 + (size_t)ELL * G * 4;
  if (smem > 200 * 1024) return false;
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
  const uint64_t ngroups = (S + G - 1) / G;
#define PVW_FUSED_ND(N)                                                                                                          \
  case N: {                                                                                                                      \
    auto kern = decode_fused_kernel<ELL, G, SW, N>;                                                                              \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;         \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem) != cudaSuccess || per_sm < 1) return false;      \
    const unsigned grid = (unsigned)std::min<uint64_t>(ngroups, (uint64_t)sms * per_sm);                                         \
    kern<<<grid, 128, smem, st>>>(z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, T, F, fb_list, fb_count);                         \
    return true;                                                                                                                 \
  }
  switch (F.nd) {
    PVW_FUSED_ND(1)
    PVW_FUSED_ND(2)
    PVW_FUSED_ND(3)
    PVW_FUSED_ND(4)
  }
#undef PVW_FUSED_ND
  return false;
}

template <int ELL, int G>
static bool launch_fused_sw(const DevTables& T, const FusedConst& F, const u64* z, size_t z_ls, size_t z_ds, size_t z_cs, const DecodeSub& sb, uint32_t Pc,
                            uint64_t S, u64* out, size_t out_ps, uint32_t* fb_list, uint32_t* fb_count, cudaStream_t st) {
  switch (T.shortSW) {
    case 2: return launch_fused_nd<ELL, G, 2>(T, F, z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, fb_list, fb_count, st);
    case 3: return launch_fused_nd<ELL, G, 3>(T, F, z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, fb_list, fb_count, st);
    case 4: return launch_fused_nd<ELL, G, 4>(T, F, z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, fb_list, fb_count, st);
  }
  return false;
}

bool launch_decode_fused(const DevTables& T, const FusedConst& F, const u64* z, size_t z_ls, size_t z_ds, uint32_t Pc, uint32_t D, u64* out, size_t out_ps,
                         uint32_t* fb_list, uint32_t* fb_count, cudaStream_t st, size_t z_cs, const DecodeSub* sub, u64* scr) {
  const uint64_t S = (uint64_t)Pc * D;
  if (S == 0) return true;
  if (!F.enabled || S >= (1ull << 32)) return false;
  const DecodeSub sb = sub ? *sub : DecodeSub{nullptr, 0, 0, nullptr, nullptr};
  switch (T.ell) {
    case 8: return launch_fused_claim_sw<8, PVW_FUSED_MINB8>(T, F, z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, fb_list, fb_count, st, scr);
    case 16: return launch_fused_claim_sw<16, 2>(T, F, z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, fb_list, fb_count, st, scr);
    case 32: return launch_fused_sw<32, 8>(T, F, z, z_ls, z_ds, z_cs, sb, Pc, S, out, out_ps, fb_list, fb_count, st);
  }
  return false;
}

void launch_decode_tail(const DevTables& T, const u64* X, uint32_t Pc, uint32_t D, u64* out, size_t out_ps, cudaStream_t st, const FallbackList* fbl) {
  const uint64_t S = (uint64_t)Pc * D;
  if (S == 0) return;
  const FallbackList fb = fbl ? *fbl : FallbackList{nullptr, nullptr};
  const unsigned blocks = share_blocks(S, 128, fb);
  if (T.tail_impl != 0) {  // register-resident specialisations for the 128- and 256-bit parameter shapes
    if (T.NW == 17 && T.divM_n == 15 && T.div2D_n == 3) { decode_tail_fixed_kernel<17, 15, 3><<<blocks, 128, 0, st>>>(X, S, Pc, T.ell, out, out_ps, T, fb); return; }
    if (T.NW == 33 && T.divM_n == 31 && T.div2D_n == 3) { decode_tail_fixed_kernel<33, 31, 3><<<blocks, 128, 0, st>>>(X, S, Pc, T.ell, out, out_ps, T, fb); return; }
    if (T.NW == 4 && T.divM_n == 4 && T.div2D_n == 1) { decode_tail_fixed_kernel<4, 4, 1><<<blocks, 128, 0, st>>>(X, S, Pc, T.ell, out, out_ps, T, fb); return; }
  }
  decode_tail_kernel<<<blocks, 128, 0, st>>>(X, S, Pc, T.ell, out, out_ps, T, fb);
}

size_t decode_fused_scratch_words(const DevTables& T, uint64_t S) { return (size_t)(T.ell + 2) * S; }
size_t decode_scratch_words_y(const DevTables& T, uint64_t S) { return (size_t)T.L * (T.ell + 1) * S; }
size_t decode_scratch_words_X(const DevTables& T, uint64_t S) { return (size_t)(T.ell + 1) * T.NW * S; }

}  // namespace pvw
