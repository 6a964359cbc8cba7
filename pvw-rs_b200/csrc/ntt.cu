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

template <int ELL, int MODE>  // MODE 0: store canonical, 1: accumulate into canonical, 2: store packed halves (operand form),
                               // 3 / 4: store the byte planes the tensor-core path reads (imma.cuh), M side / dealer side,
                               // 5: store canonical value + addend, the addend in the slot-major form the tensor-core product writes:
                               //    addend[(vec*L + limb)*ELL*inner + c*inner + j]   (unit stride across the threads of a warp)
__global__ void __launch_bounds__(128) ntt_small_kernel(const long long* __restrict__ coef, const u64* __restrict__ m, uint64_t count,
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
  const longlong2* src = reinterpret_cast<const longlong2*>(coef + idx * ELL);
#pragma unroll
  for (int t = 0; t < ELL / 2; t++) {
    longlong2 v = src[t];
    a[2 * t] = reduce_i64(v.x, lc);
    a[2 * t + 1] = reduce_i64(v.y, lc);
  }
  ntt_forward_regs<ELL>(a, s_tw, s_tw_sh, lc.q);
  if (m != nullptr) {
    u64 mr = reduce_i64((long long)m[idx], lc);  // `scalars[p] as i64`, encryption.rs:195
#pragma unroll
    for (int t = 0; t < ELL; t++) a[t] = addmod(a[t], mulmod_shoup(mr, s_g[t], s_g_sh[t], lc.q), lc.q);
  }
  if (MODE == 3 || MODE == 4) {
    // idx = r*inner + j (inner = k): r is a matrix row (MODE 3) or a dealer (MODE 4); vstride = kp, lstride = plane stride in bytes.
    // Consecutive threads hold consecutive j: every byte store of a warp fills one sector of one plane.
    const uint64_t r = idx / inner, j = idx % inner, kp = vstride;
    uint8_t* o8 = reinterpret_cast<uint8_t*>(out) + (size_t)limb * ELL * lstride;
    const size_t first = MODE == 3 ? (size_t)r * 8 * kp + j : (size_t)r * kp + j;        // byte plane 0
    const size_t step = MODE == 3 ? kp : (size_t)(count / inner) * kp;                  // to the next byte plane
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

void launch_ntt_small(const DevTables& T, const long long* coef, const u64* m, uint64_t count, uint32_t inner, u64* out,
                      size_t vstride, size_t lstride, cudaStream_t st, bool accumulate, bool pack_out, int planes, const u64* addend) {
  if (count == 0) return;
  const unsigned grid = (unsigned)(((count + 127) / 128) * T.L);
#define PVW_NTT_CASE(E)                                                                                                       \
  case E:                                                                                                                     \
    if (addend)                                                                                                               \
      ntt_small_kernel<E, 5><<<grid, 128, 0, st>>>(coef, m, count, inner, out, vstride, lstride, T.lc, T.tw, T.tw_sh, T.gadget_hat, T.gadget_hat_sh, addend, T.L);  \
    else if (planes == 1)                                                                                                     \
      ntt_small_kernel<E, 3><<<grid, 128, 0, st>>>(coef, m, count, inner, out, vstride, lstride, T.lc, T.tw, T.tw_sh, T.gadget_hat, T.gadget_hat_sh, addend, T.L);  \
    else if (planes == 2)                                                                                                     \
      ntt_small_kernel<E, 4><<<grid, 128, 0, st>>>(coef, m, count, inner, out, vstride, lstride, T.lc, T.tw, T.tw_sh, T.gadget_hat, T.gadget_hat_sh, addend, T.L);  \
    else if (accumulate)                                                                                                           \
      ntt_small_kernel<E, 1><<<grid, 128, 0, st>>>(coef, m, count, inner, out, vstride, lstride, T.lc, T.tw, T.tw_sh, T.gadget_hat, T.gadget_hat_sh, addend, T.L);  \
    else if (pack_out)                                                                                                        \
      ntt_small_kernel<E, 2><<<grid, 128, 0, st>>>(coef, m, count, inner, out, vstride, lstride, T.lc, T.tw, T.tw_sh, T.gadget_hat, T.gadget_hat_sh, addend, T.L);  \
    else                                                                                                                      \
      ntt_small_kernel<E, 0><<<grid, 128, 0, st>>>(coef, m, count, inner, out, vstride, lstride, T.lc, T.tw, T.tw_sh, T.gadget_hat, T.gadget_hat_sh, addend, T.L);  \
    break;
  switch (T.ell) {
    PVW_NTT_CASE(8)
    PVW_NTT_CASE(16)
    PVW_NTT_CASE(32)
  }
#undef PVW_NTT_CASE
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
