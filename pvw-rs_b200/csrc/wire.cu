// wire.cu -- device-side serialisation of polynomial records.
//
// Replaces the per-polynomial `Poly::to_bytes` / `Poly::from_bytes` loops of the reference's serde impls
// (src/crypto/encryption.rs:298-354, src/keys/public_key.rs:471-622, src/params/crs.rs:228-295): a ciphertext batch, rows
// of the global public key or the CRS are turned into / parsed from their wire bytes where they live, in HBM, so that only
// the (3 % smaller) wire bytes cross PCIe and no host core touches a residue.  Pure byte shuffling: HBM bound.
//
// One CTA handles up to 8 consecutive polynomials of one batch entry: their residues are staged in shared memory
// (coalesced: in the limb-major device layout the polynomials of one limb are contiguous), then every thread
// produces output bytes (pack) or residues (unpack) from shared memory, consecutive threads touching consecutive bytes.
#include <algorithm>

#include "kernels.cuh"

namespace pvw {

namespace {

constexpr uint32_t WIRE_G_MAX = 8;        // polynomials per CTA (fewer when the parameter set is too big for shared memory)
constexpr size_t WIRE_SMEM_MAX = 96 * 1024;  // keeps at least two CTAs per SM
constexpr uint32_t WIRE_THREADS = 256;

// Pack: (1) residues -> shared memory (128-bit loads); (2) one thread per (record, limb) runs the bit packer over the
// limb's ell residues and stores the bytes into a shared image of the output span; (3) the image is copied out with
// aligned 128-bit stores.  The image starts at the same offset inside a 16-byte word as the span does in global memory.
__global__ void __launch_bounds__(WIRE_THREADS) wire_pack_kernel(const WireTables W, const u64* __restrict__ src, size_t ls, size_t bs,
                                                                 uint64_t count, int src_packed, uint8_t* __restrict__ out, size_t obs, uint32_t G) {
  extern __shared__ __align__(16) u64 sv[];  // [g][L][ell] canonical residues, then the byte image
  const uint64_t p0 = (uint64_t)blockIdx.x * G;
  const uint32_t gn = (uint32_t)min((uint64_t)G, count - p0), ell = W.ell, L = W.L, pe = gn * ell, rec = W.rec_bytes;
  src += (size_t)blockIdx.y * bs + p0 * ell;
  out += (size_t)blockIdx.y * obs + p0 * rec;
  const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15);
  uint8_t* img = reinterpret_cast<uint8_t*>(sv + G * L * ell) + phase;
  const uint32_t lg_ell = 31 - __clz(ell), lg_pe = 31 - __clz(pe);
  const bool pow2 = (pe & (pe - 1)) == 0;   // true except in the last CTA of a row
  for (uint32_t idx = 2 * threadIdx.x; idx < L * pe; idx += 2 * WIRE_THREADS) {
    const uint32_t j = pow2 ? idx >> lg_pe : idx / pe, r = idx - j * pe, g = r >> lg_ell, c = r & (ell - 1);
    ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + (size_t)j * ls + r);
    if (src_packed) { v.x = unpack_halves(v.x); v.y = unpack_halves(v.y); }
    *reinterpret_cast<ulonglong2*>(sv + (g * L + j) * ell + c) = v;
  }
  for (uint32_t b = threadIdx.x; b < gn * W.pre_len; b += WIRE_THREADS) {
    const uint32_t g = b / W.pre_len, p = b - g * W.pre_len;
    img[g * rec + p] = W.pre[p];
  }
  __syncthreads();
  for (uint32_t gj = threadIdx.x; gj < gn * L; gj += WIRE_THREADS) {
    const uint32_t g = gj / L, j = gj - g * L, nb = W.nbits[j];
    const u64* row = sv + gj * ell;
    uint8_t* d = img + g * rec + W.pre_len + W.limb_off[j];
    // fhe-util transcode_to_bytes: append nb bits per residue, emit whole bytes (ell * nb is a multiple of 8)
    u64 acc = 0;
    uint32_t have = 0;
    for (uint32_t i = 0; i < ell; i++) {
      const u64 v = row[i];
      const u64 lo = acc | (v << have);                    // have < 8 pending bits, then nb <= 62 new ones: up to 69
      const u64 hi = have ? v >> (64 - have) : 0;          // the bits that did not fit
      const uint32_t tot = have + nb, nby = tot >> 3;      // whole bytes to emit: at most 8, all inside lo
      for (uint32_t t = 0; t < nby; t++) d[t] = (uint8_t)(lo >> (8 * t));
      d += nby;
      acc = nby == 8 ? hi : lo >> (8 * nby);
      have = tot & 7;
    }
  }
  __syncthreads();
  const uint32_t total = gn * rec;
  const uint32_t head = min(total, (16 - phase) & 15);
  if (threadIdx.x < head) out[threadIdx.x] = img[threadIdx.x];
  const uint32_t vecs = (total - head) / 16;
  uint4* out16 = reinterpret_cast<uint4*>(out + head);
  const uint4* img16 = reinterpret_cast<const uint4*>(img + head);
  for (uint32_t w = threadIdx.x; w < vecs; w += WIRE_THREADS) out16[w] = img16[w];
  const uint32_t tail0 = head + 16 * vecs;
  if (threadIdx.x < total - tail0) out[tail0 + threadIdx.x] = img[tail0 + threadIdx.x];
}

// Unpack: (1) the byte span -> shared image (aligned 128-bit loads); (2) header bytes are compared with the template and
// one thread per residue extracts its bits, checks them against the prime and stores them in the device layout (a
// thread-per-limb bit unpacker, the mirror of the pack kernel, measured slower: its byte loads form one long dependent chain).
__global__ void __launch_bounds__(WIRE_THREADS) wire_unpack_kernel(const WireTables W, const uint8_t* __restrict__ in, size_t ibs, uint64_t count,
                                                                   u64* __restrict__ dst, size_t ls, size_t bs, int dst_packed, int write,
                                                                   int* __restrict__ err, uint32_t G) {
  extern __shared__ __align__(16) u64 sv[];  // the byte image
  const uint64_t p0 = (uint64_t)blockIdx.x * G;
  const uint32_t gn = (uint32_t)min((uint64_t)G, count - p0), ell = W.ell, L = W.L, pe = gn * ell, rec = W.rec_bytes;
  in += (size_t)blockIdx.y * ibs + p0 * rec;
  dst += (size_t)blockIdx.y * bs + p0 * ell;
  const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 15);
  uint8_t* img = reinterpret_cast<uint8_t*>(sv) + phase;
  const uint32_t total = gn * rec;
  const uint32_t head = min(total, (16 - phase) & 15);
  if (threadIdx.x < head) img[threadIdx.x] = in[threadIdx.x];
  const uint32_t vecs = (total - head) / 16;
  const uint4* in16 = reinterpret_cast<const uint4*>(in + head);
  uint4* img16 = reinterpret_cast<uint4*>(img + head);
  for (uint32_t w = threadIdx.x; w < vecs; w += WIRE_THREADS) img16[w] = in16[w];
  const uint32_t tail0 = head + 16 * vecs;
  if (threadIdx.x < total - tail0) img[tail0 + threadIdx.x] = in[tail0 + threadIdx.x];
  __syncthreads();
  int bad = 0;
  for (uint32_t b = threadIdx.x; b < gn * W.pre_len; b += WIRE_THREADS) {
    const uint32_t g = b / W.pre_len, p = b - g * W.pre_len;
    if (img[g * rec + p] != W.pre[p]) bad |= WIRE_ERR_HEADER;
  }
  // one thread per residue: its nb bits start `sh` bits into the aligned 32-bit word holding their first byte, so at most
  // 96 bits are touched (sh <= 31, nb <= 62); the image is padded so that the third word always exists
  const uint32_t lg_ell = 31 - __clz(ell), lg_pe = 31 - __clz(pe);
  const bool pow2 = (pe & (pe - 1)) == 0;
  const uint32_t* imgw = reinterpret_cast<const uint32_t*>(img - phase);
  for (uint32_t idx = threadIdx.x; idx < L * pe; idx += WIRE_THREADS) {
    const uint32_t j = pow2 ? idx >> lg_pe : idx / pe, r = idx - j * pe, g = r >> lg_ell, c = r & (ell - 1);
    const uint32_t nb = W.nbits[j], bit = c * nb;
    const uint32_t a = phase + g * rec + W.pre_len + W.limb_off[j] + (bit >> 3);
    const uint32_t* sw = imgw + (a >> 2);
    const uint32_t sh = ((a & 3) << 3) + (bit & 7);
    const uint32_t w0 = sw[0], w1 = sw[1], w2 = sw[2];
    const u64 lo = (u64)w1 << 32 | w0;
    u64 v = sh ? (lo >> sh) | ((u64)w2 << (64 - sh)) : lo;
    v &= ~0ull >> (64 - nb);
    if (v >= W.moduli[j]) bad |= WIRE_ERR_RESIDUE;
    if (write) dst[(size_t)j * ls + r] = dst_packed ? pack_halves(v) : v;
  }
  if (bad) atomicOr(err, bad);
}

__global__ void wire_fill_kernel(uint8_t* out, size_t obs, uint32_t batch, const uint8_t* tmpl, uint32_t len) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint64_t)batch * len) return;
  const uint64_t b = i / len;
  const uint32_t o = (uint32_t)(i - b * len);
  out[b * obs + o] = tmpl[o];
}
__global__ void wire_expect_kernel(const uint8_t* in, size_t ibs, uint32_t batch, const uint8_t* tmpl, uint32_t len, int* err) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint64_t)batch * len) return;
  const uint64_t b = i / len;
  const uint32_t o = (uint32_t)(i - b * len);
  if (in[b * ibs + o] != tmpl[o]) atomicOr(err, (int)WIRE_ERR_ENVELOPE);
}

}  // namespace

static size_t pack_smem(const WireTables& W, uint32_t G) { return (size_t)G * W.L * W.ell * 8 + ((size_t)G * W.rec_bytes + 16 + 15) / 16 * 16; }
static size_t unpack_smem(const WireTables& W, uint32_t G) { return ((size_t)G * W.rec_bytes + 16 + 12 + 15) / 16 * 16; }
template <class F>
static uint32_t pick_group(const WireTables& W, F smem_of) {
  uint32_t G = WIRE_G_MAX;
  while (G > 1 && smem_of(W, G) > WIRE_SMEM_MAX) G /= 2;
  return G;
}

void launch_wire_pack(const WireTables& W, const u64* src, size_t ls, size_t bs, uint64_t count, uint32_t batch, int src_packed, uint8_t* out,
                      size_t obs, cudaStream_t st) {
  if (count == 0 || batch == 0) return;
  static const bool attr = (cudaFuncSetAttribute(wire_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), true);
  (void)attr;
  const uint32_t G = pick_group(W, pack_smem);
  for (uint32_t b0 = 0; b0 < batch; b0 += 65535) {
    dim3 grid((unsigned)((count + G - 1) / G), std::min(65535u, batch - b0));
    wire_pack_kernel<<<grid, WIRE_THREADS, pack_smem(W, G), st>>>(W, src + (size_t)b0 * bs, ls, bs, count, src_packed, out + (size_t)b0 * obs, obs, G);
  }
}
void launch_wire_unpack(const WireTables& W, const uint8_t* in, size_t ibs, uint64_t count, uint32_t batch, u64* dst, size_t ls, size_t bs,
                        int dst_packed, int write, int* err, cudaStream_t st) {
  if (count == 0 || batch == 0) return;
  static const bool attr = (cudaFuncSetAttribute(wire_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), true);
  (void)attr;
  const uint32_t G = pick_group(W, unpack_smem);
  for (uint32_t b0 = 0; b0 < batch; b0 += 65535) {
    dim3 grid((unsigned)((count + G - 1) / G), std::min(65535u, batch - b0));
    wire_unpack_kernel<<<grid, WIRE_THREADS, unpack_smem(W, G), st>>>(W, in + (size_t)b0 * ibs, ibs, count, dst + (size_t)b0 * bs, ls, bs, dst_packed, write, err, G);
  }
}
void launch_wire_fill(uint8_t* out, size_t obs, uint32_t batch, const uint8_t* tmpl, uint32_t len, cudaStream_t st) {
  if (batch == 0 || len == 0) return;
  const uint64_t total = (uint64_t)batch * len;
  wire_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, obs, batch, tmpl, len);
}
void launch_wire_expect(const uint8_t* in, size_t ibs, uint32_t batch, const uint8_t* tmpl, uint32_t len, int* err, cudaStream_t st) {
  if (batch == 0 || len == 0) return;
  const uint64_t total = (uint64_t)batch * len;
  wire_expect_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, ibs, batch, tmpl, len, err);
}

}  // namespace pvw
