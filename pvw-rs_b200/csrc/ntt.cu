// ntt.cu -- K1: per-RNS-limb ell-point negacyclic NTT held entirely in registers, fused with the RNS reduction of
// the small signed inputs and (for c2) the message encoding; plus the strided block copy used for layout changes.
//
// Replaces, on the device: Poly::from_coefficients + change_representation(Ntt) (src/crypto/encryption.rs:147-154,
// src/keys/secret_key.rs:98-112), sample_error_1/2's bigints_to_poly + NTT (src/params/parameters.rs:264-284,
// 420-474) and encode_scalar (parameters.rs:346-367).  Convention (fhe-math NttOperator): natural
// order in, bit-reversed order out, slot i = evaluation at psi^(2*brv(i)+1).
#include "kernels.cuh"
#include "ntt_regs.cuh"

namespace pvw {

// Small signed inputs (secrets, randomness, errors) arrive as int64 (the reference's i64 / BigInt-as-i64) or, to cut the
// host-to-device volume of the host-pointer calls, as int8 / int16 / int32 (pvw_b200.h PVW_IN_*): `cbytes` is the element size.
template <int N>
PVW_DEV void load_small(const void* __restrict__ coef, int cbytes, uint64_t idx, long long (&x)[N]) {
  if (cbytes == 8) {
    const longlong2* src = reinterpret_cast<const longlong2*>(reinterpret_cast<const long long*>(coef) + idx * N);
#pragma unroll
    for (int t = 0; t < N / 2; t++) { const longlong2 v = src[t]; x[2 * t] = v.x; x[2 * t + 1] = v.y; }
  } else if (cbytes == 4) {
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const int*>(coef) + idx * N);
#pragma unroll
    for (int t = 0; t < N / 4; t++) { const int4 v = src[t]; x[4 * t] = v.x; x[4 * t + 1] = v.y; x[4 * t + 2] = v.z; x[4 * t + 3] = v.w; }
  } else if (cbytes == 2) {
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const short*>(coef) + idx * N);
#pragma unroll
    for (int t = 0; t < N / 8; t++) {
      const uint4 v = src[t];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; i++) { x[8 * t + 2 * i] = (short)(w[i] & 0xffffu); x[8 * t + 2 * i + 1] = (short)(w[i] >> 16); }
    }
  } else {
    const uint2* src = reinterpret_cast<const uint2*>(reinterpret_cast<const signed char*>(coef) + idx * N);
#pragma unroll
    for (int t = 0; t < N / 8; t++) {
      const uint2 v = src[t];
      const uint32_t w[2] = {v.x, v.y};
#pragma unroll
      for (int i = 0; i < 8; i++) x[8 * t + i] = (signed char)((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
    }
  }
}

// bytes s of four residues -> eight words (word s = byte s of v[0..3]): the unit the byte-plane writers store
PVW_DEV void bytes_4x8(const u64 (&v)[4], uint32_t (&w)[8]) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t a0 = (uint32_t)(v[0] >> (32 * h)), a1 = (uint32_t)(v[1] >> (32 * h)), a2 = (uint32_t)(v[2] >> (32 * h)), a3 = (uint32_t)(v[3] >> (32 * h));
    const uint32_t t01 = __byte_perm(a0, a1, 0x5140), t23 = __byte_perm(a2, a3, 0x5140);
    const uint32_t u01 = __byte_perm(a0, a1, 0x7362), u23 = __byte_perm(a2, a3, 0x7362);
    w[4 * h + 0] = __byte_perm(t01, t23, 0x5410);
    w[4 * h + 1] = __byte_perm(t01, t23, 0x7632);
    w[4 * h + 2] = __byte_perm(u01, u23, 0x5410);
    w[4 * h + 3] = __byte_perm(u01, u23, 0x7632);
  }
}

template <int ELL, int MODE>  // MODE 0: store canonical, 1: accumulate into canonical, 2: store packed halves (operand form),
                               // 3 / 4: store the byte planes the tensor-core path reads (imma.cuh), M side / dealer side, one
                               //        polynomial per thread and single-byte stores (shapes the 4-polynomial kernel below cannot take),
                               // 5: store canonical value + addend, the addend in the slot-major form the tensor-core product writes:
                               //    addend[(vec*L + limb)*ELL*inner + c*inner + j]   (unit stride across the threads of a warp)
__global__ void __launch_bounds__(128) ntt_small_kernel(const void* __restrict__ coef, int cbytes, const u64* __restrict__ m, uint64_t count,
                                                        uint32_t inner, u64* __restrict__ out, size_t vstride, size_t lstride,
                                                        const LimbConst* __restrict__ lcs, const u64* __restrict__ tw,
                                                        const u64* __restrict__ tw_sh, const u64* __restrict__ gadget_hat,
                                                        const u64* __restrict__ gadget_hat_sh, const u64* __restrict__ addend, const uint32_t L) {
  __shared__ u64 s_tw[ELL], s_tw_sh[ELL], s_g[ELL], s_g_sh[ELL];
  // 1-D grid, limb fastest: the L CTAs that transform the same 128 polynomials run together, so the small-integer input (and
  // the message) is fetched from DRAM once, not once per limb
  const uint32_t limb = blockIdx.x % L, blk = blockIdx.x / L;
  if (threadIdx.x < ELL) {
    s_tw[threadIdx.x] = tw[(size_t)limb * ELL + threadIdx.x];
    s_tw_sh[threadIdx.x] = tw_sh[(size_t)limb * ELL + threadIdx.x];
    s_g[threadIdx.x] = gadget_hat[(size_t)limb * ELL + threadIdx.x];
    s_g_sh[threadIdx.x] = gadget_hat_sh[(size_t)limb * ELL + threadIdx.x];
  }
  __syncthreads();
  const LimbConst lc = lcs[limb];
  uint64_t idx = (uint64_t)blk * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  u64 a[ELL];
  {
    long long x[ELL];
    load_small<ELL>(coef, cbytes, idx, x);
#pragma unroll
    for (int t = 0; t < ELL; t++) a[t] = reduce_i64(x[t], lc);
  }
  ntt_forward_lazy_regs<ELL>(a, s_tw, s_tw_sh, lc.q);
  if (m != nullptr) {
    u64 mr = reduce_i64((long long)m[idx], lc);  // `scalars[p] as i64`, encryption.rs:195
#pragma unroll
    for (int t = 0; t < ELL; t++) a[t] = addmod(a[t], mulmod_shoup(mr, s_g[t], s_g_sh[t], lc.q), lc.q);
  }
  if (MODE == 3 || MODE == 4) {
    // idx = r*inner + j (inner = k): r is a matrix row (MODE 3) or a dealer (MODE 4); vstride = kp, lstride = plane stride in bytes.
    const uint64_t r = idx / inner, j = idx % inner, kp = vstride;
    uint8_t* o8 = reinterpret_cast<uint8_t*>(out) + (size_t)limb * ELL * lstride;
    const size_t first = (size_t)r * 8 * kp + j;                                        // byte plane 0 (both sides share the layout)
    const size_t step = kp;                                                             // to the next byte plane
#pragma unroll
    for (int t = 0; t < ELL; t++) {
      uint8_t* o = o8 + (size_t)t * lstride + first;
#pragma unroll
      for (int b = 0; b < 8; b++) o[(size_t)b * step] = (uint8_t)(a[t] >> (8 * b));
    }
    return;
  }
  uint64_t vec = idx / inner, j = idx % inner;
  ulonglong2* dst = reinterpret_cast<ulonglong2*>(out + vec * vstride + (size_t)limb * lstride + j * ELL);
  if (MODE == 5) {
    const u64* ad = addend + ((size_t)vec * L + limb) * ELL * inner + j;
#pragma unroll
    for (int t = 0; t < ELL; t++) a[t] = addmod(a[t], ad[(size_t)t * inner], lc.q);
  }
#pragma unroll
  for (int t = 0; t < ELL / 2; t++) {
    if (MODE == 1) {  // out += value (the matrix product was stored first; host-pointer encrypt overlaps the e2 / m copy with it)
      const ulonglong2 o = dst[t];
      dst[t] = make_ulonglong2(addmod(a[2 * t], o.x, lc.q), addmod(a[2 * t + 1], o.y, lc.q));
    } else if (MODE == 2) {
      dst[t] = make_ulonglong2(pack_halves(a[2 * t]), pack_halves(a[2 * t + 1]));
    } else {
      dst[t] = make_ulonglong2(a[2 * t], a[2 * t + 1]);
    }
  }
}

// Byte planes of NTT(small polynomial), FOUR consecutive polynomials j per thread (k % 4 == 0): every (slot, byte index) of the
// four goes out as one 4-byte word and a warp stores 128 contiguous bytes per instruction.  (The first version stored single
// bytes: 64 store instructions of 32 bytes per polynomial and limb.)  SIDE 3: matrix-row side, SIDE 4: dealer side (imma.cuh).
constexpr uint32_t TERN_PITCH = 10;     // ring degree 8: words between table rows in shared memory (8 + 2: rows start in different bank groups)
constexpr uint32_t TERN_PITCH16 = 18;   // ring degree 16
template <int ELL, int SIDE>
__global__ void __launch_bounds__(128, ELL == 8 ? 5 : ELL == 16 ? 3 : 1) ntt_planes4_kernel(const void* __restrict__ coef, int cbytes, uint64_t count, uint32_t inner, uint8_t* __restrict__ out,
                                                          size_t kp, size_t pstride, const LimbConst* __restrict__ lcs, const u64* __restrict__ tw,
                                                          const u64* __restrict__ tw_sh, const uint32_t L, const u64* __restrict__ tern,
                                                          const void* __restrict__ wide, const uint32_t* __restrict__ wide_flag) {
  __shared__ u64 s_tw[ELL], s_tw_sh[ELL];
  // `coef` may be a one-byte copy of 64-bit inputs made by narrow_i64_kernel (an eighth of the bytes for the L limbs to fetch):
  // when that kernel met a value outside [-128, 127] it raised the flag, and the original goes in instead
  if (wide_flag != nullptr && *wide_flag != 0) { coef = wide; cbytes = 8; }
  // tern != nullptr (one-byte inputs, ring degree 8): this limb's 2 x 81 transforms of ternary half-polynomials, rows TERN_PITCH words
  // apart (80 bytes: eight consecutive rows start in eight different 16-byte bank groups)
  extern __shared__ __align__(16) u64 s_tern[];
  const uint32_t limb = blockIdx.x % L, blk = blockIdx.x / L;
  if (threadIdx.x < ELL) {
    s_tw[threadIdx.x] = tw[(size_t)limb * ELL + threadIdx.x];
    s_tw_sh[threadIdx.x] = tw_sh[(size_t)limb * ELL + threadIdx.x];
  }
  if (ELL == 8 && tern != nullptr)
    for (uint32_t i = threadIdx.x; i < 162 * 8; i += blockDim.x) s_tern[(i >> 3) * TERN_PITCH + (i & 7)] = tern[(size_t)limb * 162 * 8 + i];
  if (ELL == 16 && tern != nullptr)
    for (uint32_t i = threadIdx.x; i < 324 * 16; i += blockDim.x) s_tern[(i >> 4) * TERN_PITCH16 + (i & 15)] = tern[(size_t)limb * 324 * 16 + i];
  __syncthreads();
  const LimbConst lc = lcs[limb];
  const uint64_t idx = 4 * ((uint64_t)blk * blockDim.x + threadIdx.x);   // first of this thread's four polynomials (same row: inner % 4 == 0)
  if (idx >= count) return;
  const uint64_t r = idx / inner, j = idx % inner;
  const size_t first = (size_t)r * 8 * kp + j;      // SIDE 3 and 4 share the layout: row / dealer r, byte plane b, polynomial j
  const size_t step = kp;
  uint8_t* o8 = out + (size_t)limb * ELL * pstride + first;
  if constexpr (ELL == 8) {
    // The four transforms run in a ROLLED loop (one copy of the butterflies: the straight-line form was 4 100 instructions executed
    // once per warp and a fifth of its stall cycles waited for instruction fetch); byte b of slot t of polynomial p is dropped into
    // byte p of word w[t][b] with one PRMT whose selector depends on p only.
    uint32_t w[ELL][8];
#pragma unroll
    for (int t = 0; t < ELL; t++)
#pragma unroll
      for (int b = 0; b < 8; b++) w[t][b] = 0;
#pragma unroll 1
    for (int p = 0; p < 4; p++) {
      u64 a[ELL];
      bool done = false;
      long long x[ELL];
      if (tern == nullptr || cbytes != 1) load_small<ELL>(coef, cbytes, idx + p, x);
      if (tern != nullptr) {
        // Secrets and encryption randomness are ternary at the reference's default variance (CBD, parameters.rs:166,251-254): the
        // transform is linear, so NTT(x) = T_lo[x_0..x_3] + T_hi[x_4..x_7] -- two table rows and eight modular additions instead
        // of twelve butterflies.  Any other coefficient takes the butterflies below (same canonical result).
        uint32_t ilo = 0, ihi = 0;
        bool ternary;
        if (cbytes == 1) {
          const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const signed char*>(coef) + (idx + p) * 8);
          const uint32_t t0 = ((v.x & 0x7f7f7f7fu) + 0x01010101u) ^ (v.x & 0x80808080u);    // x_i + 1 in every byte (no carries across bytes)
          const uint32_t t1 = ((v.y & 0x7f7f7f7fu) + 0x01010101u) ^ (v.y & 0x80808080u);
          ternary = (((t0 | t1) & 0xfcfcfcfcu) | (((t0 & (t0 >> 1)) | (t1 & (t1 >> 1))) & 0x01010101u)) == 0;   // every byte in {0, 1, 2}
          ilo = __dp4a(t0, 0x1b090301u, 0u);
          ihi = __dp4a(t1, 0x1b090301u, 0u);
          if (!ternary) load_small<ELL>(coef, cbytes, idx + p, x);
        } else {
          u64 worst = 0;
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const u64 u0 = (u64)x[t] + 1, u1 = (u64)x[4 + t] + 1;
            worst |= u0 | u1;                                                               // {0, 1, 2} or'ed together stay below 4
            ilo = ilo + (uint32_t)u0 * (t == 0 ? 1u : t == 1 ? 3u : t == 2 ? 9u : 27u);
            ihi = ihi + (uint32_t)u1 * (t == 0 ? 1u : t == 1 ? 3u : t == 2 ? 9u : 27u);
          }
          ternary = worst < 4;
#pragma unroll
          for (int t = 0; t < ELL; t++) ternary = ternary && (u64)x[t] + 1 != 3;
        }
        if (ternary) {
          const ulonglong2* lo = reinterpret_cast<const ulonglong2*>(s_tern + (size_t)ilo * TERN_PITCH);
          const ulonglong2* hi = reinterpret_cast<const ulonglong2*>(s_tern + (size_t)(81u + ihi) * TERN_PITCH);
#pragma unroll
          for (int t = 0; t < ELL / 2; t++) {
            const ulonglong2 l = lo[t], h = hi[t];
            a[2 * t] = addmod(l.x, h.x, lc.q);
            a[2 * t + 1] = addmod(l.y, h.y, lc.q);
          }
          done = true;
        }
      }
      if (!done) {
#pragma unroll
        for (int t = 0; t < ELL; t++) a[t] = reduce_i64(x[t], lc);
        ntt_forward_lazy_regs<ELL>(a, s_tw, s_tw_sh, lc.q);
      }
      const uint32_t keep = 0x3210u & ~(0xFu << (4 * p));                  // every byte of w but byte p
      uint32_t sel[4];
#pragma unroll
      for (int b = 0; b < 4; b++) sel[b] = keep | ((4u + b) << (4 * p));   // byte p <- byte b of the source word
#pragma unroll
      for (int t = 0; t < ELL; t++) {
        const uint32_t lo = (uint32_t)a[t], hi = (uint32_t)(a[t] >> 32);
#pragma unroll
        for (int b = 0; b < 4; b++) { w[t][b] = __byte_perm(w[t][b], lo, sel[b]); w[t][4 + b] = __byte_perm(w[t][4 + b], hi, sel[b]); }
      }
    }
#pragma unroll
    for (int t = 0; t < ELL; t++)
#pragma unroll
      for (int b = 0; b < 8; b++) *reinterpret_cast<uint32_t*>(o8 + (size_t)t * pstride + (size_t)b * step) = w[t][b];
  } else if (ELL == 16 && tern != nullptr) {
    // Ring degree 16 with the ternary tables: 8 * 16 output words do not fit the registers, so the slots go out in two halves; a
    // ternary polynomial costs four table rows (one per group of four coefficients) and three modular additions per slot in each
    // half, any other polynomial a full transform per half (rare: secrets and randomness at the default variance are ternary).
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      uint32_t w[8][8];
#pragma unroll
      for (int t = 0; t < 8; t++)
#pragma unroll
        for (int b = 0; b < 8; b++) w[t][b] = 0;
#pragma unroll 1
      for (int p = 0; p < 4; p++) {
        u64 a[8];
        uint32_t gi[4] = {0, 0, 0, 0};
        bool ternary;
        if (cbytes == 1) {
          const uint2* src = reinterpret_cast<const uint2*>(reinterpret_cast<const signed char*>(coef) + (idx + p) * 16);   // (8-byte loads: the
          const uint2 v0 = src[0], v1 = src[1];                                                                            //  alignment load_small asks for)
          const uint32_t vw[4] = {v0.x, v0.y, v1.x, v1.y};
          uint32_t bad = 0;
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const uint32_t t = ((vw[g] & 0x7f7f7f7fu) + 0x01010101u) ^ (vw[g] & 0x80808080u);   // x_i + 1 in every byte
            bad |= (t & 0xfcfcfcfcu) | ((t & (t >> 1)) & 0x01010101u);
            gi[g] = __dp4a(t, 0x1b090301u, 0u);
          }
          ternary = bad == 0;
        } else {
          long long x[ELL];
          load_small<ELL>(coef, cbytes, idx + p, x);
          ternary = true;
#pragma unroll
          for (int i = 0; i < ELL; i++) {
            const u64 u = (u64)x[i] + 1;
            ternary = ternary && u <= 2;
            gi[i >> 2] += (uint32_t)u * ((i & 3) == 0 ? 1u : (i & 3) == 1 ? 3u : (i & 3) == 2 ? 9u : 27u);
          }
        }
        if (ternary) {
          const u64* r0 = s_tern + (size_t)gi[0] * TERN_PITCH16 + 8 * half;
          const u64* r1 = s_tern + (size_t)(81u + gi[1]) * TERN_PITCH16 + 8 * half;
          const u64* r2 = s_tern + (size_t)(162u + gi[2]) * TERN_PITCH16 + 8 * half;
          const u64* r3 = s_tern + (size_t)(243u + gi[3]) * TERN_PITCH16 + 8 * half;
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const ulonglong2 v0 = reinterpret_cast<const ulonglong2*>(r0)[t], v1 = reinterpret_cast<const ulonglong2*>(r1)[t];
            const ulonglong2 v2 = reinterpret_cast<const ulonglong2*>(r2)[t], v3 = reinterpret_cast<const ulonglong2*>(r3)[t];
            a[2 * t] = addmod(addmod(v0.x, v1.x, lc.q), addmod(v2.x, v3.x, lc.q), lc.q);
            a[2 * t + 1] = addmod(addmod(v0.y, v1.y, lc.q), addmod(v2.y, v3.y, lc.q), lc.q);
          }
        } else {
          long long x[ELL];
          load_small<ELL>(coef, cbytes, idx + p, x);
          u64 f[ELL];
#pragma unroll
          for (int t = 0; t < ELL; t++) f[t] = reduce_i64(x[t], lc);
          ntt_forward_lazy_regs<ELL>(f, s_tw, s_tw_sh, lc.q);
#pragma unroll
          for (int t = 0; t < 8; t++) a[t] = half ? f[8 + t] : f[t];
        }
        const uint32_t keep = 0x3210u & ~(0xFu << (4 * p));
        uint32_t sel[4];
#pragma unroll
        for (int b = 0; b < 4; b++) sel[b] = keep | ((4u + b) << (4 * p));
#pragma unroll
        for (int t = 0; t < 8; t++) {
          const uint32_t lo = (uint32_t)a[t], hi = (uint32_t)(a[t] >> 32);
#pragma unroll
          for (int b = 0; b < 4; b++) { w[t][b] = __byte_perm(w[t][b], lo, sel[b]); w[t][4 + b] = __byte_perm(w[t][4 + b], hi, sel[b]); }
        }
      }
#pragma unroll
      for (int t = 0; t < 8; t++)
#pragma unroll
        for (int b = 0; b < 8; b++) *reinterpret_cast<uint32_t*>(o8 + (size_t)(8 * half + t) * pstride + (size_t)b * step) = w[t][b];
    }
  } else {   // 8 * ELL output words do not fit the registers next to a transform: straight-line, four transforms side by side
    u64 a[4][ELL];
#pragma unroll
    for (int p = 0; p < 4; p++) {
      long long x[ELL];
      load_small<ELL>(coef, cbytes, idx + p, x);
#pragma unroll
      for (int t = 0; t < ELL; t++) a[p][t] = reduce_i64(x[t], lc);
      ntt_forward_lazy_regs<ELL>(a[p], s_tw, s_tw_sh, lc.q);
    }
#pragma unroll
    for (int t = 0; t < ELL; t++) {
      const u64 v[4] = {a[0][t], a[1][t], a[2][t], a[3][t]};
      uint32_t w[8];
      bytes_4x8(v, w);
#pragma unroll
      for (int b = 0; b < 8; b++) *reinterpret_cast<uint32_t*>(o8 + (size_t)t * pstride + (size_t)b * step) = w[b];
    }
  }
}

// c1 = NTT(e1) + A r_hat, finished where both halves meet: the tensor-core product arrives slot-major (unit stride across the
// threads), the result leaves twice -- as packed-halves residues into the ciphertext store (operand of the CUDA-core kernels,
// source of downloads and of the wire format) and as byte planes right behind them in the same store slot (operand of the
// tensor-core decryption, and what the multi-GPU exchange ships), so that decryption needs no conversion pass over c1.
// Four polynomials (rows j of c1) per thread; k % 4 == 0.
template <int ELL>
__global__ void __launch_bounds__(128) ntt_c1_finish_kernel(const void* __restrict__ coef, int cbytes, uint64_t count, uint32_t k, u64* __restrict__ c1,
                                                            size_t slot_stride, const u64* __restrict__ addend, uint32_t kp,
                                                            const LimbConst* __restrict__ lcs, const u64* __restrict__ tw, const u64* __restrict__ tw_sh,
                                                            const uint32_t L) {
  __shared__ u64 s_tw[ELL], s_tw_sh[ELL];
  const uint32_t limb = blockIdx.x % L, blk = blockIdx.x / L;
  if (threadIdx.x < ELL) {
    s_tw[threadIdx.x] = tw[(size_t)limb * ELL + threadIdx.x];
    s_tw_sh[threadIdx.x] = tw_sh[(size_t)limb * ELL + threadIdx.x];
  }
  __syncthreads();
  const LimbConst lc = lcs[limb];
  const uint64_t idx = 4 * ((uint64_t)blk * blockDim.x + threadIdx.x);
  if (idx >= count) return;
  const uint64_t d = idx / k, j = idx % k;
  u64 a[4][ELL];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    long long x[ELL];
    load_small<ELL>(coef, cbytes, idx + p, x);
#pragma unroll
    for (int t = 0; t < ELL; t++) a[p][t] = reduce_i64(x[t], lc);
    ntt_forward_lazy_regs<ELL>(a[p], s_tw, s_tw_sh, lc.q);
  }
  const u64* ad = addend + ((size_t)d * L + limb) * ELL * k + j;
#pragma unroll
  for (int t = 0; t < ELL; t++) {
    const ulonglong2 v0 = *reinterpret_cast<const ulonglong2*>(ad + (size_t)t * k), v1 = *reinterpret_cast<const ulonglong2*>(ad + (size_t)t * k + 2);
    a[0][t] = addmod(a[0][t], v0.x, lc.q); a[1][t] = addmod(a[1][t], v0.y, lc.q);
    a[2][t] = addmod(a[2][t], v1.x, lc.q); a[3][t] = addmod(a[3][t], v1.y, lc.q);
  }
  u64* slot = c1 + d * slot_stride;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(slot + (size_t)limb * k * ELL + (j + p) * ELL);
#pragma unroll
    for (int t = 0; t < ELL / 2; t++) dst[t] = make_ulonglong2(pack_halves(a[p][2 * t]), pack_halves(a[p][2 * t + 1]));
  }
  uint8_t* o8 = reinterpret_cast<uint8_t*>(slot + (size_t)L * k * ELL) + (size_t)limb * ELL * 8 * kp + j;   // planes [L*ELL][8][kp] behind the residues
#pragma unroll
  for (int t = 0; t < ELL; t++) {
    const u64 v[4] = {a[0][t], a[1][t], a[2][t], a[3][t]};
    uint32_t w[8];
    bytes_4x8(v, w);
#pragma unroll
    for (int b8 = 0; b8 < 8; b8++) *reinterpret_cast<uint32_t*>(o8 + ((size_t)t * 8 + b8) * kp) = w[b8];
  }
}

// 64-bit small inputs -> one byte each (eight per thread); *flag is raised when a value does not fit, and the consumer then reads
// the original (ntt_planes4_kernel).  Secrets are ternary or a few units wide (CBD, parameters.rs:251-254), the reference keeps them as i64.
__global__ void __launch_bounds__(256) narrow_i64_kernel(const long long* __restrict__ in, signed char* __restrict__ out, uint64_t groups, uint32_t* __restrict__ flag) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups) return;
  const longlong2* src = reinterpret_cast<const longlong2*>(in + i * 8);
  uint32_t w[2] = {0, 0};
  bool bad = false;
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const longlong2 v = src[t];
    bad = bad || v.x < -128 || v.x > 127 || v.y < -128 || v.y > 127;
    w[t >> 1] |= ((uint32_t)v.x & 0xffu) << (16 * (t & 1)) | ((uint32_t)v.y & 0xffu) << (16 * (t & 1) + 8);
  }
  *reinterpret_cast<uint2*>(out + i * 8) = make_uint2(w[0], w[1]);
  if (bad) *flag = 1u;
}

bool launch_narrow_i64(const void* in, void* out, uint64_t values, uint32_t* flag, cudaStream_t st) {
  if (values == 0) return true;
  if (values % 8 != 0) return false;
  const uint64_t groups = values / 8, blocks = (groups + 255) / 256;
  if (blocks >= (1ull << 31)) return false;
  narrow_i64_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(in), reinterpret_cast<signed char*>(out), groups, flag);
  return true;
}

bool launch_ntt_c1_finish(const DevTables& T, const void* coef, int cbytes, uint64_t count, uint32_t k, u64* c1, size_t slot_stride, const u64* addend,
                          uint32_t kp, cudaStream_t st) {
  if (count == 0) return true;
  if (k % 4 != 0 || (T.ell != 8 && T.ell != 16) || (cbytes != 1 && cbytes != 2 && cbytes != 4 && cbytes != 8)) return false;
  const uint64_t blocks = ((count / 4 + 127) / 128) * T.L;
  if (blocks >= (1ull << 31)) return false;
  if (T.ell == 8) ntt_c1_finish_kernel<8><<<(unsigned)blocks, 128, 0, st>>>(coef, cbytes, count, k, c1, slot_stride, addend, kp, T.lc, T.tw, T.tw_sh, T.L);
  else ntt_c1_finish_kernel<16><<<(unsigned)blocks, 128, 0, st>>>(coef, cbytes, count, k, c1, slot_stride, addend, kp, T.lc, T.tw, T.tw_sh, T.L);
  return true;
}

// Any power-of-two ring degree up to GEN_MAX_ELL (the reference accepts every power of two >= 8, parameters.rs:140-144; its
// tests and examples use 8, 16, 32): the same transform with the coefficients in a per-thread local-memory array and run-time
// loops.  Correct for every mode, not tuned -- the register-resident kernels above serve the parameter sets in use.
constexpr int GEN_MAX_ELL = 256;
__global__ void __launch_bounds__(64) ntt_small_generic_kernel(const int mode, const void* __restrict__ coef, int cbytes, const u64* __restrict__ m,
                                                               uint64_t count, uint32_t inner, u64* __restrict__ out, size_t vstride, size_t lstride,
                                                               const LimbConst* __restrict__ lcs, const u64* __restrict__ tw, const u64* __restrict__ tw_sh,
                                                               const u64* __restrict__ gadget_hat, const u64* __restrict__ gadget_hat_sh,
                                                               const u64* __restrict__ addend, const uint32_t L, const uint32_t ell) {
  const uint32_t limb = blockIdx.x % L, blk = blockIdx.x / L;
  const LimbConst lc = lcs[limb];
  const uint64_t idx = (uint64_t)blk * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  u64 a[GEN_MAX_ELL];
  for (uint32_t t = 0; t < ell; t++) {
    long long x;
    const uint64_t e = idx * ell + t;
    if (cbytes == 8) x = reinterpret_cast<const long long*>(coef)[e];
    else if (cbytes == 4) x = reinterpret_cast<const int*>(coef)[e];
    else if (cbytes == 2) x = reinterpret_cast<const short*>(coef)[e];
    else x = reinterpret_cast<const signed char*>(coef)[e];
    a[t] = reduce_i64(x, lc);
  }
  const u64* w = tw + (size_t)limb * ell;
  const u64* w_sh = tw_sh + (size_t)limb * ell;
  for (uint32_t mm = 1, t = ell >> 1; mm < ell; mm <<= 1, t >>= 1)
    for (uint32_t i = 0; i < mm; i++) {
      const u64 s = w[mm + i], s_sh = w_sh[mm + i];
      for (uint32_t j = 2 * i * t; j < 2 * i * t + t; j++) {
        const u64 u = a[j], v = mulmod_shoup(a[j + t], s, s_sh, lc.q);
        a[j] = addmod(u, v, lc.q);
        a[j + t] = submod(u, v, lc.q);
      }
    }
  if (m != nullptr) {
    const u64 mr = reduce_i64((long long)m[idx], lc);
    for (uint32_t t = 0; t < ell; t++)
      a[t] = addmod(a[t], mulmod_shoup(mr, gadget_hat[(size_t)limb * ell + t], gadget_hat_sh[(size_t)limb * ell + t], lc.q), lc.q);
  }
  if (mode == 3 || mode == 4) {
    const uint64_t r = idx / inner, j = idx % inner, kp = vstride;
    uint8_t* o8 = reinterpret_cast<uint8_t*>(out) + (size_t)limb * ell * lstride;
    const size_t first = (size_t)r * 8 * kp + j;
    const size_t step = kp;
    for (uint32_t t = 0; t < ell; t++)
      for (int b = 0; b < 8; b++) o8[(size_t)t * lstride + first + (size_t)b * step] = (uint8_t)(a[t] >> (8 * b));
    return;
  }
  const uint64_t vec = idx / inner, j = idx % inner;
  u64* dst = out + vec * vstride + (size_t)limb * lstride + j * ell;
  for (uint32_t t = 0; t < ell; t++) {
    u64 v = a[t];
    if (mode == 5) v = addmod(v, addend[((size_t)vec * L + limb) * ell * inner + (size_t)t * inner + j], lc.q);
    if (mode == 1) v = addmod(v, dst[t], lc.q);
    dst[t] = mode == 2 ? pack_halves(v) : v;
  }
}

bool launch_ntt_small(const DevTables& T, const void* coef, int cbytes, const u64* m, uint64_t count, uint32_t inner, u64* out,
                      size_t vstride, size_t lstride, cudaStream_t st, bool accumulate, bool pack_out, int planes, const u64* addend,
                      const void* wide, const uint32_t* wide_flag) {
  if (count == 0) return true;
  if (cbytes != 1 && cbytes != 2 && cbytes != 4 && cbytes != 8) return false;
  const int mode = addend ? 5 : planes == 1 ? 3 : planes == 2 ? 4 : accumulate ? 1 : pack_out ? 2 : 0;
  const uint64_t blocks = ((count + 127) / 128) * T.L;
  if (blocks >= (1ull << 31)) return false;
  const unsigned grid = (unsigned)blocks;
  // byte planes, four polynomials per thread: needs whole groups of four inside a row and registers for 4 * ell residues
  const bool four = (mode == 3 || mode == 4) && inner % 4 == 0 && T.ell <= 16 && m == nullptr;
  const unsigned grid4 = (unsigned)(((count / 4 + 127) / 128) * T.L);
  if (wide_flag != nullptr && !four) return false;          // a narrowed copy is only understood by the four-polynomial byte-plane kernels
  // ring degree 8: the table path for ternary polynomials (secrets, randomness); T.tern is null when switched off
  const u64* tern = (four && (T.ell == 8 || T.ell == 16)) ? T.tern : nullptr;
  const size_t tern_smem = !tern ? 0 : T.ell == 8 ? (size_t)162 * TERN_PITCH * 8 : (size_t)324 * TERN_PITCH16 * 8;
#define PVW_NTT_ARGS coef, cbytes, m, count, inner, out, vstride, lstride, T.lc, T.tw, T.tw_sh, T.gadget_hat, T.gadget_hat_sh, addend, T.L
#define PVW_NTT_CASE(E)                                                                                                          \
  case E:                                                                                                                        \
    if (four && mode == 3) ntt_planes4_kernel<(E <= 16 ? E : 8), 3><<<grid4, 128, tern_smem, st>>>(coef, cbytes, count, inner, reinterpret_cast<uint8_t*>(out), vstride, lstride, T.lc, T.tw, T.tw_sh, T.L, tern, wide, wide_flag); \
    else if (four) ntt_planes4_kernel<(E <= 16 ? E : 8), 4><<<grid4, 128, tern_smem, st>>>(coef, cbytes, count, inner, reinterpret_cast<uint8_t*>(out), vstride, lstride, T.lc, T.tw, T.tw_sh, T.L, tern, wide, wide_flag);        \
    else if (mode == 5) ntt_small_kernel<E, 5><<<grid, 128, 0, st>>>(PVW_NTT_ARGS);                                              \
    else if (mode == 3) ntt_small_kernel<E, 3><<<grid, 128, 0, st>>>(PVW_NTT_ARGS);                                              \
    else if (mode == 4) ntt_small_kernel<E, 4><<<grid, 128, 0, st>>>(PVW_NTT_ARGS);                                              \
    else if (mode == 1) ntt_small_kernel<E, 1><<<grid, 128, 0, st>>>(PVW_NTT_ARGS);                                              \
    else if (mode == 2) ntt_small_kernel<E, 2><<<grid, 128, 0, st>>>(PVW_NTT_ARGS);                                              \
    else ntt_small_kernel<E, 0><<<grid, 128, 0, st>>>(PVW_NTT_ARGS);                                                             \
    break;
  switch (T.ell) {
    PVW_NTT_CASE(8)
    PVW_NTT_CASE(16)
    PVW_NTT_CASE(32)
    default: {
      if (T.ell > (uint32_t)GEN_MAX_ELL) return false;
      const uint64_t gb = ((count + 63) / 64) * T.L;
      if (gb >= (1ull << 31)) return false;
      ntt_small_generic_kernel<<<(unsigned)gb, 64, 0, st>>>(mode, PVW_NTT_ARGS, T.ell);
    }
  }
#undef PVW_NTT_CASE
#undef PVW_NTT_ARGS
  return true;
}

// out[b*obs + x*oxs + y*oys + c] = in[b*ibs + x*ixs + y*iys + c]; one thread per 16 bytes (blk is a multiple of 2)
template <int XFORM>
__global__ void __launch_bounds__(256) permute_kernel(const ulonglong2* __restrict__ in, ulonglong2* __restrict__ out, uint64_t total,
                                                      uint64_t X, uint64_t Y, uint32_t blk2, size_t ibs, size_t ixs, size_t iys, size_t obs,
                                                      size_t oxs, size_t oys) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  uint32_t c = (uint32_t)(e % blk2);
  uint64_t r = e / blk2;
  uint64_t x = r % X; r /= X;   // x fastest: consecutive threads walk the output's contiguous axis when oxs == blk
  uint64_t y = r % Y;
  uint64_t b = r / Y;
  ulonglong2 v = in[(b * ibs + x * ixs + y * iys) / 2 + c];
  if (XFORM == 1) v = make_ulonglong2(pack_halves(v.x), pack_halves(v.y));
  if (XFORM == 2) v = make_ulonglong2(unpack_halves(v.x), unpack_halves(v.y));
  out[(b * obs + x * oxs + y * oys) / 2 + c] = v;
}

void launch_permute(const u64* in, u64* out, uint64_t Bn, uint64_t X, uint64_t Y, uint32_t blk, size_t ibs, size_t ixs, size_t iys,
                    size_t obs, size_t oxs, size_t oys, cudaStream_t st, int xform) {
  uint64_t total = Bn * X * Y * (blk / 2);
  if (total == 0) return;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  const ulonglong2* i2 = reinterpret_cast<const ulonglong2*>(in);
  ulonglong2* o2 = reinterpret_cast<ulonglong2*>(out);
  if (xform == 1) permute_kernel<1><<<blocks, 256, 0, st>>>(i2, o2, total, X, Y, blk / 2, ibs, ixs, iys, obs, oxs, oys);
  else if (xform == 2) permute_kernel<2><<<blocks, 256, 0, st>>>(i2, o2, total, X, Y, blk / 2, ibs, ixs, iys, obs, oxs, oys);
  else permute_kernel<0><<<blocks, 256, 0, st>>>(i2, o2, total, X, Y, blk / 2, ibs, ixs, iys, obs, oxs, oys);
}

}  // namespace pvw
