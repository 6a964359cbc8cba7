// mac.cu -- K2/K3/K5: the NTT-domain polynomial matrix product that carries >95 % of the hot path's arithmetic.
//
//   acc[row][d][c] = sum_{j<k} M[limb][row][j][c] * V[d][limb][j][c]   (mod q_limb), one CTA = one limb x (RT rows x DT dealers)
//
// One kernel serves four reference loops (same arithmetic, different operands):
//   c1 = A r + e1      PvwCrs::multiply_by_randomness, src/params/crs.rs:177-205 + src/crypto/encryption.rs:161-173
//   c2 = B r + e2 + mg the per-party loop,             src/crypto/encryption.rs:177-200
//   z  = <s, c1> - c2  decrypt_party_value,            src/crypto/decryption.rs:257-274
//   b  = A^T s + e     multiply_by_secret_key/keygen,  src/params/crs.rs:138-171, src/keys/public_key.rs:111-147
//
// B200 mapping.  The reference calls this once per (dealer, party); batched over D dealers it is a modular GEMM per
// slot, so each M row tile fetched from HBM/L2 is reused for DT dealers and each V tile for RT rows.  The bound then
// is the integer pipe, not HBM: a 62x62-bit product is 4 IMAD.WIDE.U32; products are accumulated unreduced in 160
// bits (modarith.cuh) and reduced once per k terms.  No tensor cores: exact modular integer arithmetic.
//   * thread = one NTT slot c of a TR x TD (row, dealer) sub-tile -> TR*TD independent carry chains (ILP);
//     a warp's lanes sweep the ell slots of (32/ell) sub-tiles, so shared-memory reads are conflict-free 8-byte
//     accesses, broadcast across the sub-tiles that share a row.
//   * operand rows are contiguous in the limb-major layout (kc*ell*8 bytes): a producer warp streams them with
//     cp.async.bulk (TMA, SASS UBLKCP) into a 4-stage shared-memory ring guarded by mbarriers (impl 1);
//     impl 0 is the same tile with synchronous loads (bring-up / cross-check path).
//   * D == 1 (a single `encrypt` call) degenerates to the HBM-bound matrix-vector product with DT = 1.
#include <cstdio>

#include "kernels.cuh"

namespace pvw {

constexpr int kComputeThreads = 256;
constexpr int NS = 4;  // pipeline stages

template <int ELL, int TR, int TD, int GD, int KC>
struct TileCfg {
  // KC = polynomials (j indices) per pipeline stage
  static constexpr int G = kComputeThreads / ELL;  // (row-group, dealer-group) pairs per CTA
  static constexpr int GR = G / GD;
  static constexpr int RT = GR * TR;               // rows per CTA
  static constexpr int DT = GD * TD;               // dealers per CTA
  static constexpr int ROWB = KC * ELL * 8 + 16;   // bytes per staged row (+16: dealer sub-tiles land on distinct banks)
  static constexpr int STAGE = (RT + DT) * ROWB;
  static_assert(G % GD == 0, "bad tile");
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();  // never hang the device: a lost barrier becomes a launch error
  } while (!ok);
}
// TMA 1D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int ELL, int TR, int TD, int GD, int KC>
struct Worker {
  using C = TileCfg<ELL, TR, TD, GD, KC>;
  int c, gr, gd;
  Acc160 acc[TR][TD];
  __device__ __forceinline__ void init(int tid) {
    const int lane = tid & 31, w = tid >> 5;
    c = lane % ELL;
    const int g = w * (32 / ELL) + lane / ELL;
    gd = g % GD;
    gr = g / GD;
#pragma unroll
    for (int t = 0; t < TR; t++)
#pragma unroll
      for (int u = 0; u < TD; u++) acc_zero(acc[t][u]);
  }
  // one staged chunk: kc polynomials of every row / dealer of the tile
  template <bool FULL>
  __device__ __forceinline__ void chunk(const unsigned char* stage, int kc) {
    const unsigned char* ms = stage + (size_t)(gr * TR) * C::ROWB + c * 8;
    const unsigned char* vs = stage + (size_t)(C::RT + gd * TD) * C::ROWB + c * 8;
#pragma unroll
    for (int jj = 0; jj < KC; jj++) {
      if (!FULL && jj >= kc) break;
      u64 a[TR], b[TD];
#pragma unroll
      for (int t = 0; t < TR; t++) a[t] = *reinterpret_cast<const u64*>(ms + t * C::ROWB + jj * ELL * 8);
#pragma unroll
      for (int u = 0; u < TD; u++) b[u] = *reinterpret_cast<const u64*>(vs + u * C::ROWB + jj * ELL * 8);
#pragma unroll
      for (int t = 0; t < TR; t++)
#pragma unroll
        for (int u = 0; u < TD; u++) acc_mac(acc[t][u], a[t], b[u]);
    }
  }
  __device__ __forceinline__ void epilogue(const GemmArgs& g, uint32_t limb, uint32_t r0, uint32_t d0) {
    const LimbConst lc = g.lc[limb];
#pragma unroll
    for (int u = 0; u < TD; u++) {
      const uint32_t d = d0 + gd * TD + u;
      if (d >= g.D) continue;
      const uint32_t ds = g.V_dmap ? g.V_dmap[d] : d;
#pragma unroll
      for (int t = 0; t < TR; t++) {
        const uint32_t row = r0 + gr * TR + t;
        if (row >= g.rows) continue;
        u64 v = acc_reduce(acc[t][u], lc);
        u64* o = g.O + (size_t)d * g.O_ds + (size_t)limb * g.O_ls + (size_t)row * ELL + c;
        if (g.mode == 0) {
          v = addmod(v, *o, lc.q);
        } else {
          const uint32_t srow = g.S_rowmap ? g.S_rowmap[row] : row;
          v = submod(v, g.S[(size_t)ds * g.S_ds + (size_t)limb * g.S_ls + (size_t)srow * ELL + c], lc.q);
        }
        *o = v;
      }
    }
  }
};

// source address of staged row `rr` of the tile (rows first, then dealers); out-of-range indices are clamped to the
// last valid one (their products are computed and discarded), so every copy is in bounds and full width.
template <class C>
__device__ __forceinline__ const u64* row_src(const GemmArgs& g, uint32_t limb, uint32_t r0, uint32_t d0, int rr, int ELL) {
  if (rr < C::RT) {
    uint32_t row = min(r0 + (uint32_t)rr, g.rows - 1);
    return g.M + (size_t)limb * g.M_ls + (size_t)row * g.M_rs;
  }
  uint32_t d = min(d0 + (uint32_t)(rr - C::RT), g.D - 1);
  if (g.V_dmap) d = g.V_dmap[d];
  return g.V + (size_t)d * g.V_ds + (size_t)limb * g.V_ls;
}

// ---- impl 0: synchronous tiles ------------------------------------------------------------------------------------
template <int ELL, int TR, int TD, int GD, int KC>
__global__ void __launch_bounds__(kComputeThreads, 1) mac_gemm_sync_kernel(const GemmArgs g) {
  using C = TileCfg<ELL, TR, TD, GD, KC>;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t limb = blockIdx.z, r0 = blockIdx.x * C::RT, d0 = blockIdx.y * C::DT;
  Worker<ELL, TR, TD, GD, KC> wk;
  wk.init(threadIdx.x);
  for (uint32_t j0 = 0; j0 < g.k; j0 += KC) {
    const int kc = min((uint32_t)KC, g.k - j0);
    const int vec_per_row = kc * ELL / 2;  // 16-byte vectors
    for (int e = threadIdx.x; e < (C::RT + C::DT) * vec_per_row; e += kComputeThreads) {
      const int rr = e / vec_per_row, v = e % vec_per_row;
      const ulonglong2* src = reinterpret_cast<const ulonglong2*>(row_src<C>(g, limb, r0, d0, rr, ELL) + (size_t)j0 * ELL);
      reinterpret_cast<ulonglong2*>(smem + (size_t)rr * C::ROWB)[v] = src[v];
    }
    __syncthreads();
    if (kc == KC) wk.template chunk<true>(smem, kc); else wk.template chunk<false>(smem, kc);
    __syncthreads();
  }
  wk.epilogue(g, limb, r0, d0);
}

// ---- impl 1: TMA bulk-copy producer warp + mbarrier ring ---------------------------------------------------------
template <int ELL, int TR, int TD, int GD, int KC>
__global__ void __launch_bounds__(kComputeThreads + 32, 1) mac_gemm_tma_kernel(const GemmArgs g) {
  using C = TileCfg<ELL, TR, TD, GD, KC>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[2 * NS];  // full[NS], empty[NS]
  const uint32_t limb = blockIdx.z, r0 = blockIdx.x * C::RT, d0 = blockIdx.y * C::DT;
  const int tid = threadIdx.x;
  const uint32_t bar0 = smem_u32(bars);
  if (tid == 0) {
    for (int s = 0; s < NS; s++) {
      mbar_init(bar0 + 8 * s, 1);                            // full: one arrive.expect_tx by the producer
      mbar_init(bar0 + 8 * (NS + s), kComputeThreads / 32);  // empty: one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t nchunks = (g.k + KC - 1) / KC;
  if (tid >= kComputeThreads) {
    // ===== producer warp: one 1D bulk copy per staged row =====
    const int lane = tid & 31;
    const u64* src[(C::RT + C::DT + 31) / 32];
#pragma unroll
    for (int i = 0; i < (C::RT + C::DT + 31) / 32; i++) {
      const int rr = lane + 32 * i;
      src[i] = rr < C::RT + C::DT ? row_src<C>(g, limb, r0, d0, rr, ELL) : nullptr;
    }
    for (uint32_t it = 0; it < nchunks; it++) {
      const int s = it % NS;
      const uint32_t n = it / NS;
      mbar_wait(bar0 + 8 * (NS + s), (n & 1) ^ 1);  // consumers released the previous use of this stage
      const uint32_t kc = min((uint32_t)KC, g.k - it * KC);
      const uint32_t row_bytes = kc * ELL * 8;
      if (lane == 0) mbar_expect_tx(bar0 + 8 * s, (C::RT + C::DT) * row_bytes);
      __syncwarp();
      const uint32_t dst0 = smem_u32(smem) + s * C::STAGE;
#pragma unroll
      for (int i = 0; i < (C::RT + C::DT + 31) / 32; i++) {
        const int rr = lane + 32 * i;
        if (rr < C::RT + C::DT) bulk_g2s(dst0 + rr * C::ROWB, src[i] + (size_t)it * KC * ELL, row_bytes, bar0 + 8 * s);
      }
    }
    return;
  }
  // ===== consumer warps =====
  Worker<ELL, TR, TD, GD, KC> wk;
  wk.init(tid);
  for (uint32_t it = 0; it < nchunks; it++) {
    const int s = it % NS;
    const uint32_t n = it / NS;
    mbar_wait(bar0 + 8 * s, n & 1);
    const int kc = min((uint32_t)KC, g.k - it * KC);
    const unsigned char* stage = smem + (size_t)s * C::STAGE;
    if (kc == KC) wk.template chunk<true>(stage, kc); else wk.template chunk<false>(stage, kc);
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(bar0 + 8 * (NS + s));
  }
  wk.epilogue(g, limb, r0, d0);
}

template <int ELL, int TR, int TD, int GD, int KC>
static void launch_cfg(const GemmArgs& a, int impl, cudaStream_t st) {
  using C = TileCfg<ELL, TR, TD, GD, KC>;
  dim3 grid((a.rows + C::RT - 1) / C::RT, (a.D + C::DT - 1) / C::DT, a.L);
  if (impl == 0) {
    auto kern = mac_gemm_sync_kernel<ELL, TR, TD, GD, KC>;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::STAGE); attr = true; }
    kern<<<grid, kComputeThreads, C::STAGE, st>>>(a);
  } else {
    auto kern = mac_gemm_tma_kernel<ELL, TR, TD, GD, KC>;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, NS * C::STAGE); attr = true; }
    kern<<<grid, kComputeThreads + 32, NS * C::STAGE, st>>>(a);
  }
}

void launch_mac_gemm(const GemmArgs& a, int impl, cudaStream_t st) {
  if (a.rows == 0 || a.D == 0) return;
  const bool matvec = a.D == 1;
  switch (a.ell) {
    case 8:
      if (matvec) launch_cfg<8, 4, 1, 1, 4>(a, impl, st); else launch_cfg<8, 4, 4, 4, 8>(a, impl, st);
      break;
    case 16:
      if (matvec) launch_cfg<16, 4, 1, 1, 4>(a, impl, st); else launch_cfg<16, 4, 4, 4, 8>(a, impl, st);
      break;
    case 32:
      if (matvec) launch_cfg<32, 4, 1, 1, 4>(a, impl, st); else launch_cfg<32, 4, 4, 2, 4>(a, impl, st);
      break;
  }
}
size_t mac_gemm_launches(const GemmArgs& a) { return (a.rows == 0 || a.D == 0) ? 0 : 1; }

}  // namespace pvw
