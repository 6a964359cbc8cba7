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
// is the integer pipe, not HBM: a 62x62-bit product is 3 IMAD.WIDE.U32 (Karatsuba on 31-bit halves); the three partial
// sums are accumulated unreduced in 96 bits each (modarith.cuh, AccK) and reduced once per k terms.  Measured on
// B200 (tools/csrc/int_peaks.cu): an IMAD.WIDE with a 64-bit addend issues every 4 cycles per SM sub-partition, so the
// ceiling is 12 cycles per warp-wide multiply-accumulate = 2.8e12 MAC/s.  No tensor cores: exact modular arithmetic.
//   * thread = one NTT slot c of a TR x TD (row, dealer) sub-tile -> TR*TD independent carry chains (ILP);
//     a warp's lanes sweep the ell slots of (32/ell) sub-tiles, so shared-memory reads are conflict-free 8-byte
//     accesses, broadcast across the sub-tiles that share a row.
//   * operand rows are contiguous in the limb-major layout (kc*ell*8 bytes): the compute warps stream them with
//     cp.async.bulk (TMA, SASS UBLKCP) into a 4-stage shared-memory ring guarded by mbarriers (impl 1);
//     impl 0 is the same tile with synchronous loads (bring-up / cross-check path).
//   * D == 1 (a single `encrypt` call) degenerates to the HBM-bound matrix-vector product with DT = 1.
#include <cuda.h>

#include <cstdio>

#include "kernels.cuh"
#include "mac_worker.cuh"

namespace pvw {

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
template <int ELL, int TR, int TD, int GD, int KC, int NB>
__global__ void __launch_bounds__(kComputeThreads, NB) mac_gemm_sync_kernel(const GemmArgs g) {
  using C = TileCfg<ELL, TR, TD, GD, KC>;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t limb = blockIdx.z, r0 = blockIdx.x * C::RT, d0 = blockIdx.y * C::DT;
  Worker<ELL, TR, TD, GD, KC> wk;
  wk.init(threadIdx.x, (u32)g.lc[limb].pad);
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

// ---- impl 1: TMA bulk copies + mbarrier ring, no dedicated producer warp ---------------------------------------------
// Nine warps would cap the kernel at 168 registers (three warps on one sub-partition share 16 K registers), so the
// eight compute warps feed themselves: warp w owns rows {w, w+8, ...} of the staged tile; before it starts chunk i it
// refills the stage that chunk i-1 used (everybody has left it: `empty` barrier) with chunk i-1+NS.
template <int ELL, int TR, int TD, int GD, int KC, int NB, int NS>
__global__ void __launch_bounds__(kComputeThreads, NB) mac_gemm_tma_kernel(const GemmArgs g) {
  using C = TileCfg<ELL, TR, TD, GD, KC>;
  constexpr int NW = kComputeThreads / 32;
  constexpr int ROWS = C::RT + C::DT;
  constexpr int RPW = (ROWS + NW - 1) / NW;  // staged rows per warp
  static_assert(RPW <= 32, "one lane per staged row");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[2 * NS];  // full[NS], empty[NS]
  const uint32_t limb = blockIdx.z, r0 = blockIdx.x * C::RT, d0 = blockIdx.y * C::DT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar0 = smem_u32(bars);
  if (tid == 0) {
    for (int s = 0; s < NS; s++) {
      mbar_init(bar0 + 8 * s, NW);         // full: one arrive.expect_tx per warp (its rows' bytes)
      mbar_init(bar0 + 8 * (NS + s), NW);  // empty: one arrive per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t nchunks = (g.k + KC - 1) / KC;
  // this lane's staged row (rows first, then dealers)
  const int my_row = warp + NW * lane;
  const bool issuer = lane < RPW && my_row < ROWS;
  const int my_nrows = (ROWS - warp + NW - 1) / NW;  // rows owned by this warp
  const u64* my_src = issuer ? row_src<C>(g, limb, r0, d0, my_row, ELL) : nullptr;
  auto refill = [&](uint32_t chunk) {
    const int s = chunk % NS;
    const uint32_t kc = min((uint32_t)KC, g.k - chunk * KC);
    const uint32_t row_bytes = kc * ELL * 8;
    if (lane == 0) mbar_expect_tx(bar0 + 8 * s, my_nrows * row_bytes);
    __syncwarp();
    if (issuer) bulk_g2s(smem_u32(smem) + s * C::STAGE + my_row * C::ROWB, my_src + (size_t)chunk * KC * ELL, row_bytes, bar0 + 8 * s);
  };
  for (uint32_t ch = 0; ch < nchunks && ch < (uint32_t)NS; ch++) refill(ch);
  // the stage of chunk i-lag is refilled when chunk i starts: lag 1 = deepest prefetch but the wait on `empty` acts as
  // a per-chunk barrier between the warps; lag 2 leaves a chunk of slack
  const uint32_t lag = g.refill_lag >= 1 && g.refill_lag < NS ? g.refill_lag : 1;
  Worker<ELL, TR, TD, GD, KC> wk;
  wk.init(tid, (u32)g.lc[limb].pad);
  for (uint32_t it = 0; it < nchunks; it++) {
    if (it >= lag && it - lag + NS < nchunks) {
      const uint32_t prev = it - lag;
      mbar_wait(bar0 + 8 * (NS + prev % NS), (prev / NS) & 1);  // every warp has finished chunk it-lag
      refill(prev + NS);
    }
    const int s = it % NS;
    mbar_wait(bar0 + 8 * s, (it / NS) & 1);
    const int kc = min((uint32_t)KC, g.k - it * KC);
    const unsigned char* stage = smem + (size_t)s * C::STAGE;
    if (kc == KC) wk.template chunk<true>(stage, kc); else wk.template chunk<false>(stage, kc);
    __syncwarp();
    if (lane == 0) mbar_arrive(bar0 + 8 * (NS + s));
  }
  wk.epilogue(g, limb, r0, d0);
}

// ---- impl 2: the matrix tile as ONE tensor-map TMA box per chunk ---------------------------------------------------------
// cp.async.bulk.tensor.3d over M viewed as u64[L][rows][k*ell]: box = KC*ell x RT x 1.  Out-of-range rows / polynomials
// are zero-filled by the TMA unit (their products vanish), so there is no tail path.  Only warp 0 feeds the ring: lane 0
// issues the box, lanes 0..DT-1 one bulk copy each for the dealer rows; the other warps just wait and compute.
__device__ __forceinline__ void tensor_g2s_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
template <int ELL, int TR, int TD, int GD, int KC, int NB, int NS>
__global__ void __launch_bounds__(kComputeThreads, NB) mac_gemm_tensor_kernel(const GemmArgs g, const __grid_constant__ CUtensorMap tmapM) {
  using W = Worker<ELL, TR, TD, GD, KC, 4, false, true, kComputeThreads, true>;
  using C = typename W::C;
  constexpr int NW = kComputeThreads / 32;
  static_assert(C::DT <= 32, "one lane per dealer row");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[2 * NS];  // full[NS], empty[NS]
  const uint32_t limb = blockIdx.z, r0 = blockIdx.x * C::RT, d0 = blockIdx.y * C::DT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar0 = smem_u32(bars);
  if (tid == 0) {
    for (int s = 0; s < NS; s++) {
      mbar_init(bar0 + 8 * s, 1);          // full: one arrive.expect_tx by the feeding lane
      mbar_init(bar0 + 8 * (NS + s), NW);  // empty: one arrive per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t nchunks = (g.k + KC - 1) / KC;
  const u64* my_v = nullptr;
  if (warp == 0 && lane < C::DT) {
    uint32_t d = min(d0 + (uint32_t)lane, g.D - 1);
    if (g.V_dmap) d = g.V_dmap[d];
    my_v = g.V + (size_t)d * g.V_ds + (size_t)limb * g.V_ls;
  }
  auto refill = [&](uint32_t chunk) {  // warp 0 only
    const int s = chunk % NS;
    const uint32_t kc = min((uint32_t)KC, g.k - chunk * KC);
    const uint32_t v_bytes = kc * ELL * 8;
    const uint32_t dst = smem_u32(smem) + s * C::STAGE;
    if (lane == 0) {
      mbar_expect_tx(bar0 + 8 * s, C::MBYTES + C::DT * v_bytes);
      tensor_g2s_3d(dst, &tmapM, (int)(chunk * KC * ELL), (int)r0, (int)limb, bar0 + 8 * s);
    }
    __syncwarp();
    if (lane < C::DT) bulk_g2s(dst + C::MBYTES + lane * C::VROWB, my_v + (size_t)chunk * KC * ELL, v_bytes, bar0 + 8 * s);
  };
  if (warp == 0)
    for (uint32_t ch = 0; ch < nchunks && ch < (uint32_t)NS; ch++) refill(ch);
  W wk;
  wk.init(tid, (u32)g.lc[limb].pad);
  const uint32_t lag = g.refill_lag >= 1 && g.refill_lag < NS ? g.refill_lag : 1;
  for (uint32_t it = 0; it < nchunks; it++) {
    if (warp == 0 && it >= lag && it - lag + NS < nchunks) {
      const uint32_t prev = it - lag;
      mbar_wait(bar0 + 8 * (NS + prev % NS), (prev / NS) & 1);  // every warp has finished chunk it-lag
      refill(prev + NS);
    }
    const int s = it % NS;
    mbar_wait(bar0 + 8 * s, (it / NS) & 1);
    wk.template chunk<true>(smem + (size_t)s * C::STAGE, KC);  // a short last chunk is zero-filled by the TMA unit
    __syncwarp();
    if (lane == 0) mbar_arrive(bar0 + 8 * (NS + s));
  }
  wk.epilogue(g, limb, r0, d0);
}

// tensor map over M as u64[L][rows][k*ell] (driver entry point fetched at run time: no link-time libcuda dependency)
static bool make_tensor_map(CUtensorMap* tm, const GemmArgs& a, uint32_t box_inner, uint32_t box_rows) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<encode_fn>(fn);
  }
  if (!encode) return false;
  if (((uintptr_t)a.M & 15) || ((a.M_rs * 8) & 15) || ((a.M_ls * 8) & 15)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)a.k * a.ell, a.rows, a.L};
  const cuuint64_t strides[2] = {(cuuint64_t)a.M_rs * 8, (cuuint64_t)a.M_ls * 8};
  const cuuint32_t box[3] = {box_inner, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<u64*>(a.M), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int ELL, int TR, int TD, int GD, int KC, int NB, int NS = kStages>
static void launch_cfg(const GemmArgs& a, int impl, cudaStream_t st) {
  using C = TileCfg<ELL, TR, TD, GD, KC>;
  dim3 grid((a.rows + C::RT - 1) / C::RT, (a.D + C::DT - 1) / C::DT, a.L);
  if (impl == 0) {
    auto kern = mac_gemm_sync_kernel<ELL, TR, TD, GD, KC, NB>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::STAGE);  // per device: set on every launch
    kern<<<grid, kComputeThreads, C::STAGE, st>>>(a);
  } else if (impl == 2 && KC * ELL <= 256) {
    using CD = TileCfg<ELL, TR, TD, GD, KC, kComputeThreads, true>;
    CUtensorMap tm;
    if (!make_tensor_map(&tm, a, KC * ELL, CD::RT)) { launch_cfg<ELL, TR, TD, GD, KC, NB, NS>(a, 1, st); return; }
    auto kern = mac_gemm_tensor_kernel<ELL, TR, TD, GD, KC, NB, NS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, NS * CD::STAGE);
    kern<<<grid, kComputeThreads, NS * CD::STAGE, st>>>(a, tm);
  } else {
    auto kern = mac_gemm_tma_kernel<ELL, TR, TD, GD, KC, NB, NS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, NS * C::STAGE);
    kern<<<grid, kComputeThreads, NS * C::STAGE, st>>>(a);
  }
}

// Any ring degree (run-time ell; used above 32, where the tiled kernels have no instantiation): one thread per output slot,
// operands read from global memory (packed halves), the same lazy accumulator and epilogue.  Correct, not tuned.
__global__ void __launch_bounds__(256) mac_gemm_generic_kernel(const GemmArgs g) {
  const uint32_t limb = blockIdx.z, d = blockIdx.y;
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (uint64_t)g.rows * g.ell) return;
  const uint32_t row = (uint32_t)(e / g.ell), c = (uint32_t)(e % g.ell);
  const LimbConst lc = g.lc[limb];
  const uint32_t ds = g.V_dmap ? g.V_dmap[d] : d;
  const u64* mrow = g.M + (size_t)limb * g.M_ls + (size_t)row * g.M_rs + c;
  const u64* vrow = g.V + (size_t)ds * g.V_ds + (size_t)limb * g.V_ls + c;
  AccK acc;
  acck_zero(acc);
  for (uint32_t j = 0; j < g.k; j++) {
    const u64 a = mrow[(size_t)j * g.ell], b = vrow[(size_t)j * g.ell];
    SplitOp sa, sb;
    sa.x0 = (u32)a; sa.x1 = (u32)(a >> 32); sa.xs = sa.x0 + sa.x1;
    sb.x0 = (u32)b; sb.x1 = (u32)(b >> 32); sb.xs = sb.x0 + sb.x1;
    acck_mac(acc, sa, sb);
  }
  u64 v = acck_reduce(acc, lc);
  u64* o = g.O + (size_t)d * g.O_ds + (size_t)limb * g.O_ls + (size_t)row * g.ell + c;
  if (g.mode == 0) v = addmod(v, *o, lc.q);
  else if (g.mode == 1) {
    const uint32_t srow = g.S_rowmap ? g.S_rowmap[row] : row;
    v = submod(v, g.S[(size_t)ds * g.S_ds + (size_t)limb * g.S_ls + (size_t)srow * g.ell + c], lc.q);
  }
  *o = g.O_packed ? pack_halves(v) : v;
}

bool launch_mac_gemm(const GemmArgs& a, int impl, cudaStream_t st) {
  if (a.rows == 0 || a.D == 0) return true;
  if (a.ell != 8 && a.ell != 16 && a.ell != 32) {
    if (a.D > 65535u || a.L > 65535u) return false;
    mac_gemm_generic_kernel<<<dim3((unsigned)(((uint64_t)a.rows * a.ell + 255) / 256), a.D, a.L), 256, 0, st>>>(a);
    return true;
  }
  const bool matvec = a.D == 1;
  // tile 1 (default): 4x2 register tile, two CTAs per SM (<= 128 registers: one CTA's prologue / epilogue overlaps the
  //   other's main loop); CTA tile 16 rows x 16 dealers (fewest staged rows per output), 16 polynomials per stage, three
  //   stages refilled with a lag of two chunks.  Measured steps (C3, ms of mac_gemm per bench step): 8 polynomials x 4
  //   stages 43.0 -> 16 x 2 39.6 (the per-chunk barrier round trip is the overhead that matters) -> 16 x 16 tile, 3 stages,
  //   lag 2: 38.5 (slack against warp skew).  tile 2: 32 x 8 tile, 8 x 4 stages; tile 0: 4x4 register tile, one CTA per SM
  const int tile = a.tile;
  switch (a.ell) {
    case 8:
      if (matvec && tile == 0) launch_cfg<8, 2, 1, 1, 8, 1>(a, impl, st);
      else if (matvec && tile == 2) launch_cfg<8, 1, 1, 1, 16, 1>(a, impl, st);
      else if (matvec) launch_cfg<8, 1, 1, 1, 16, 2, 3>(a, impl, st);   // D = 1: two CTAs/SM x 3 stages of 1 KB row copies: 91-95 % of HBM
      else if (tile == 0) launch_cfg<8, 4, 4, 4, 8, 1>(a, impl, st);
      else if (tile == 2) launch_cfg<8, 4, 2, 4, 8, 2>(a, impl, st);
      else if (tile == 3) launch_cfg<8, 4, 2, 4, 16, 2, 2>(a, impl, st);   // 32 rows x 8 dealers, 16 polynomials x 2 stages
      else launch_cfg<8, 4, 2, 8, 16, 2, 3>(a, impl, st);
      break;
    case 16:
      if (matvec) launch_cfg<16, 2, 1, 1, 8, 2, 3>(a, impl, st);
      else if (tile == 0) launch_cfg<16, 4, 4, 4, 8, 1>(a, impl, st);
      else if (tile == 2) launch_cfg<16, 4, 2, 4, 8, 2>(a, impl, st);
      else launch_cfg<16, 4, 2, 4, 16, 2, 2>(a, impl, st);
      break;
    case 32:
      if (matvec) launch_cfg<32, 4, 1, 1, 4, 1>(a, impl, st);
      else if (tile == 0) launch_cfg<32, 4, 4, 2, 4, 1>(a, impl, st);
      else if (tile == 2) launch_cfg<32, 4, 2, 2, 4, 2>(a, impl, st);
      else launch_cfg<32, 4, 2, 2, 8, 2, 2>(a, impl, st);
      break;
  }
  return true;
}
size_t mac_gemm_launches(const GemmArgs& a) { return (a.rows == 0 || a.D == 0) ? 0 : 1; }

}  // namespace pvw
