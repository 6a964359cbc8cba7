// imma.cu -- the modular matrix product on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
//   acc[row][d] = sum_{j<k} M[row][j] * V[d][j]  (mod q)      per plane = (RNS limb, NTT slot); M, V < 2^62
//
// Batched over dealers the product is a GEMM per plane and, on the CUDA cores, bound by the integer pipe (mac.cu: three
// IMAD.WIDE per 62-bit multiply-accumulate, 2.8e12 MAC/s ceiling).  Exact integer GEMMs are what the INT8 tensor cores do:
// a 64-bit operand IS its 8 little-endian bytes, so
//     M * V = sum_{u=0..14} 2^(8u) * sum_{s+t=u} m_s * v_t .
// The M row is used as it lies in memory (K axis = its 8k bytes); V is expanded once per batch into 15 "diagonal" rows
//     Vx[(d,u)][8j + s] = v_{u-s}(d, j)       (0 outside 0..7)
// and ONE u8 x u8 -> s32 GEMM  C[row][(d,u)] = sum_{(j,s)} M''[row][(j,s)] * Vx[(d,u)][(j,s)]  delivers the 15 diagonal sums
// of every output (each < 8 k 255^2 < 2^31 for k <= 4096) in the TMEM lane of its row: the epilogue thread that owns the
// lane recombines them into one 160-bit integer and reduces it -- no cross-lane traffic, no re-layout of the big operand.
// 120 int8 multiply-accumulates per 62-bit one (64 would do with both operands in byte planes, at the price of an
// 8-lane shuffle reduction of multi-word integers in the epilogue).
//
// Kernel shape (the canonical Blackwell GEMM: TMA -> 128B-swizzled smem ring -> tcgen05.mma -> TMEM -> tcgen05.ld epilogue):
//   CTA tile 256 rows x 16 dealers (240 MMA columns) x all of K; two M = 128 accumulators (2 x 256 TMEM columns) share every
//   staged Vx tile; K advances 128 bytes per stage (one swizzle span = 4 MMAs of K = 32); 3 stages of 62 KB.
//   warp 0: TMA producer, warp 1: TMEM allocation + MMA issue (one lane), warps 2-5: epilogue (one TMEM lane group each).
//   Persistent: one CTA per SM walks the tile list; the smem ring keeps filling across tile boundaries.
#include <cuda.h>

#include <algorithm>

#include "imma.cuh"

namespace pvw {

namespace {

constexpr uint32_t RT = 256, DT = 16, NT = DT * IMMA_DIAGS, KC = 128, NS = 3;
constexpr uint32_t A_BYTES = RT * KC, B_BYTES = NT * KC, B_SLOT = 32768, STAGE = A_BYTES + B_SLOT;
constexpr uint32_t THREADS = 192;
constexpr uint32_t SMEM_BYTES = NS * STAGE + 1024 /* alignment slack */ + 128 /* barriers, TMEM pointer */;
static_assert(8 * (2 * NS + 2) + 4 <= 128, "barrier block too small");
constexpr uint32_t TMEM_COLS = 512, ACC_COLS = 256;
// instruction descriptor (kind::i8): D = s32 (bits 4-5 = 2), A and B unsigned 8-bit (bits 7-9, 10-12 = 0), both K-major
// (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28
constexpr uint32_t IDESC = (2u << 4) | ((NT >> 3) << 17) | ((128u >> 4) << 24);

#define IMMA_DEV __device__ __forceinline__

IMMA_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
IMMA_DEV void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
IMMA_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
IMMA_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();  // never hang the device: a lost barrier becomes a launch error
  } while (!ok);
}
IMMA_DEV void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
IMMA_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
IMMA_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
IMMA_DEV void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32, M = 128, N = NT, K = 32
IMMA_DEV void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
// 16 consecutive columns of this thread's TMEM lane
IMMA_DEV void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}
IMMA_DEV void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (SBO), version 1 (sm_100)
IMMA_DEV uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// sum_u s[u] * 2^(8u)  mod q   (s[u] < 2^32; the true value is < k * 2^124, so the fifth word stays small)
IMMA_DEV u64 recombine(const uint32_t (&s)[16], const LimbConst& lc) {
  // diagonals u = r mod 4 are word aligned among themselves: X_r = (s[r], s[4+r], s[8+r], s[12+r]) as a 128-bit integer, and the
  // value is X_0 + (X_1 << 8) + (X_2 << 16) + (X_3 << 24)
  uint32_t w[5];
  u64 acc = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const uint32_t c0 = i < 4 ? s[4 * i] : 0u, c1 = i < 4 ? s[4 * i + 1] : 0u, c2 = i < 4 ? s[4 * i + 2] : 0u, c3 = i < 3 ? s[4 * i + 3] : 0u;
    const uint32_t p1 = i > 0 ? s[4 * i - 3] : 0u, p2 = i > 0 ? s[4 * i - 2] : 0u, p3 = (i > 0 && i < 4) ? s[4 * i - 1] : 0u;   // s[15] is the next dealer's column
    acc += (u64)c0 + __funnelshift_l(p1, c1, 8) + __funnelshift_l(p2, c2, 16) + __funnelshift_l(p3, c3, 24);
    w[i] = (uint32_t)acc;
    acc >>= 32;
  }
  const u64 lo = ((u64)w[1] << 32) | w[0], hi = ((u64)w[3] << 32) | w[2];
  const u64 h = reduce128(reduce64((u64)w[4], lc), hi, lc);
  return reduce128(h, lo, lc);
}

__global__ void __launch_bounds__(THREADS, 1) imma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                               const ImmaArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // 128-byte swizzle wants 1024-byte aligned tiles
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + NS * STAGE);        // full[NS], empty[NS], tmem_full, tmem_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2);
  const uint32_t bar0 = base + NS * STAGE;
  auto full = [&](uint32_t s) { return bar0 + 8 * s; };
  auto empty = [&](uint32_t s) { return bar0 + 8 * (NS + s); };
  const uint32_t tmem_full = bar0 + 8 * 2 * NS, tmem_empty = tmem_full + 8;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t nkb = (8 * g.k + KC - 1) / KC;
  // persistent: CTA b takes tiles b, b + gridDim.x, ...; row tiles vary fastest, then dealer tiles, then planes, so the CTAs
  // running at any moment share one plane's operands in L2.  The smem ring runs on across tiles: the next tile's first
  // stages arrive while this tile's epilogue drains TMEM.
  const uint32_t n_rt = (g.rows + RT - 1) / RT, n_dt = (g.D + DT - 1) / DT, total = n_rt * n_dt * g.L * g.ell;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < NS; s++) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, THREADS - 64);                                 // every epilogue thread arrives
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t kbg = 0;
      for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const uint32_t rt = tile % n_rt, dt = (tile / n_rt) % n_dt, plane = tile / (n_rt * n_dt);
        for (uint32_t kb = 0; kb < nkb; kb++, kbg++) {
          const uint32_t s = kbg % NS, it = kbg / NS;
          mbar_wait(empty(s), (it & 1) ^ 1);                             // first pass over the ring: passes at once
          mbar_expect_tx(full(s), A_BYTES + B_BYTES);
          tma_load_3d(base + s * STAGE, &tmA, (int)(kb * KC), (int)(rt * RT), (int)plane, full(s));
          tma_load_3d(base + s * STAGE + A_BYTES, &tmB, (int)(kb * KC), (int)(dt * DT * IMMA_DIAGS), (int)plane, full(s));
        }
      }
    }
  } else if (warp == 1) {
    uint32_t kbg = 0, i = 0;
    for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x, i++) {
      mbar_wait(tmem_empty, (i & 1) ^ 1);                                // the epilogue has read the previous tile out of TMEM
      tc_fence_after();
      for (uint32_t kb = 0; kb < nkb; kb++, kbg++) {
        const uint32_t s = kbg % NS, it = kbg / NS;
        mbar_wait(full(s), it & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = base + s * STAGE, b_addr = a_addr + A_BYTES;
#pragma unroll
          for (uint32_t k4 = 0; k4 < KC / 32; k4++) {
            const uint64_t bdesc = umma_desc(b_addr + 32 * k4);
            tc_mma_i8(tmem_base, umma_desc(a_addr + 32 * k4), bdesc, (kb | k4) != 0);
            tc_mma_i8(tmem_base + ACC_COLS, umma_desc(a_addr + 128 * KC + 32 * k4), bdesc, (kb | k4) != 0);
          }
          tc_commit(empty(s));                                           // arrives when the MMAs above have read the stage
          if (kb + 1 == nkb) tc_commit(tmem_full);
        }
        __syncwarp();
      }
    }
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w % 4) .. +31; thread = one row of each accumulator
    const uint32_t lg = warp & 3;
    uint32_t i = 0;
    for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x, i++) {
      const uint32_t rt = tile % n_rt, dt = (tile / n_rt) % n_dt, plane = tile / (n_rt * n_dt);
      const uint32_t row0 = rt * RT, d0 = dt * DT;
      const uint32_t limb = plane / g.ell, c = plane - limb * g.ell;
      const LimbConst lc = g.lc[limb];
      // the addend (mode 0: the pre-loaded O, mode 1: S) does not depend on the product: fetch all 32 of this thread's values
      // while the MMAs run, so that the epilogue proper never waits for global memory
      u64 pre[2][DT];
      size_t o_row[2];
      bool row_ok[2];
#pragma unroll
      for (uint32_t t = 0; t < 2; t++) {
        const uint32_t row = row0 + t * 128 + lg * 32 + lane;
        row_ok[t] = row < g.rows;
        o_row[t] = (size_t)limb * g.O_ls + (size_t)row * g.O_rs + (size_t)c * g.O_cs;
        const uint32_t srow = (g.mode == 1 && row_ok[t]) ? (g.S_rowmap ? g.S_rowmap[row] : row) : 0;
#pragma unroll
        for (uint32_t dd = 0; dd < DT; dd++) {
          const uint32_t d = d0 + dd;
          u64 v = 0;
          if (g.mode != 2 && row_ok[t] && d < g.D) {
            if (g.mode == 0) v = g.O[(size_t)d * g.O_ds + o_row[t]];
            else {
              const uint32_t sd = g.V_dmap ? g.V_dmap[d] : d;
              v = g.S[(size_t)sd * g.S_ds + (size_t)limb * g.S_ls + (size_t)srow * g.ell + c];
            }
          }
          pre[t][dd] = v;
        }
      }
      mbar_wait(tmem_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (uint32_t t = 0; t < 2; t++) {
#pragma unroll
        for (uint32_t dd = 0; dd < DT; dd++) {
          uint32_t s[16];
          tc_ld16(tmem_base + ((lg * 32) << 16) + t * ACC_COLS + dd * IMMA_DIAGS, s);
          tc_ld_wait();
          if (t == 1 && dd == DT - 1) {                                  // last read of this tile: hand TMEM back to the MMA warp
            tc_fence_before();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty) : "memory");
          }
          const uint32_t d = d0 + dd;
          if (row_ok[t] && d < g.D) {
            u64 r = recombine(s, lc);
            if (g.mode == 0) r = addmod(r, pre[t][dd], lc.q);
            else if (g.mode == 1) r = submod(r, pre[t][dd], lc.q);
            g.O[(size_t)d * g.O_ds + o_row[t]] = g.O_packed ? pack_halves(r) : r;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// Vx rows of one (dealer, limb): thread = (j, c) with c fastest (the read is one contiguous run)
__global__ void __launch_bounds__(256) imma_expand_kernel(const u64* __restrict__ V, size_t V_ds, size_t V_ls, uint32_t ell, uint32_t k,
                                                          uint8_t* __restrict__ Vx, size_t Vx_plane, int packed, const uint32_t* __restrict__ dmap) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y, limb = blockIdx.z;
  if (e >= k * ell) return;
  const uint32_t j = e / ell, c = e - j * ell;
  const uint32_t sd = dmap ? dmap[d] : d;
  u64 v = V[(size_t)sd * V_ds + (size_t)limb * V_ls + e];
  if (packed) v = unpack_halves(v);
  // R = v with its bytes reversed: row u of the expansion is R >> 8(7-u) for u <= 7 and R << 8(u-7) above
  const u64 R = ((u64)__byte_perm((uint32_t)v, 0, 0x0123) << 32) | __byte_perm((uint32_t)(v >> 32), 0, 0x0123);
  u64* out = reinterpret_cast<u64*>(Vx + (size_t)(limb * ell + c) * Vx_plane) + (size_t)d * IMMA_DIAGS * k + j;
#pragma unroll
  for (uint32_t u = 0; u < IMMA_DIAGS; u++) out[(size_t)u * k] = u <= 7 ? R >> (8 * (7 - u)) : R << (8 * (u - 7));
}

// one CTA per (row, limb): thread = output element (c, j), j fastest
__global__ void __launch_bounds__(256) imma_slot_major_kernel(const u64* __restrict__ M, size_t M_ls, size_t M_rs, uint32_t k, uint32_t ell,
                                                              u64* __restrict__ out, size_t out_plane, int packed) {
  const uint32_t row = blockIdx.x, limb = blockIdx.y;
  const u64* src = M + (size_t)limb * M_ls + (size_t)row * M_rs;
  for (uint32_t t = threadIdx.x; t < k * ell; t += blockDim.x) {
    const uint32_t c = t / k, j = t - c * k;
    u64 v = src[(size_t)j * ell + c];
    if (packed) v = unpack_halves(v);
    out[(size_t)(limb * ell + c) * out_plane + (size_t)row * k + j] = v;
  }
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_fn tensor_map_encoder() {
  static encode_fn encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<encode_fn>(fn);
  }
  return encode;
}
// u8 tensor [planes][nrows][kbytes], box = 128 bytes x box_rows x 1, 128-byte swizzle, out-of-range elements read as zero
bool make_map(CUtensorMap* tm, const void* base, uint64_t kbytes, uint64_t nrows, uint64_t planes, uint64_t plane_stride_bytes, uint32_t box_rows) {
  encode_fn encode = tensor_map_encoder();
  if (!encode || ((uintptr_t)base & 15) || (kbytes & 15) || (plane_stride_bytes & 15)) return false;
  const cuuint64_t dims[3] = {kbytes, nrows, planes};
  const cuuint64_t strides[2] = {kbytes, plane_stride_bytes};
  const cuuint32_t box[3] = {KC, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool imma_shape_ok(uint32_t rows, uint32_t D, uint32_t k) {
  // 8 * k * 255^2 must fit the signed 32-bit accumulator; global strides must be multiples of 16 bytes
  return rows > 0 && D > 0 && k >= 2 && k % 2 == 0 && k <= 4096;
}

bool launch_imma_gemm(const ImmaArgs& a, cudaStream_t st) {
  if (!imma_shape_ok(a.rows, a.D, a.k)) return false;
  CUtensorMap tmA, tmB;
  const uint64_t kbytes = 8ull * a.k, planes = (uint64_t)a.L * a.ell;
  if (!make_map(&tmA, a.M, kbytes, a.rows, planes, a.M_plane * 8, RT)) return false;
  if (!make_map(&tmB, a.Vx, kbytes, (uint64_t)a.D * IMMA_DIAGS, planes, a.Vx_plane, NT)) return false;
  static const bool attr = (cudaFuncSetAttribute(imma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES), true);
  (void)attr;
  cudaFuncSetAttribute(imma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);  // per device
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t tiles = (uint64_t)((a.rows + RT - 1) / RT) * ((a.D + DT - 1) / DT) * planes;
  if (tiles >= (1ull << 32)) return false;
  imma_gemm_kernel<<<(unsigned)std::min<uint64_t>(tiles, (uint64_t)std::max(sms, 1)), THREADS, SMEM_BYTES, st>>>(tmA, tmB, a);
  return true;
}

void launch_imma_expand(const u64* V, size_t V_ds, size_t V_ls, uint32_t ell, uint32_t D, uint32_t k, uint32_t L, uint8_t* Vx, size_t Vx_plane,
                        bool packed, const uint32_t* dmap, cudaStream_t st) {
  if (D == 0) return;
  dim3 grid((k * ell + 255) / 256, D, L);
  imma_expand_kernel<<<grid, 256, 0, st>>>(V, V_ds, V_ls, ell, k, Vx, Vx_plane, packed ? 1 : 0, dmap);
}

void launch_imma_slot_major(const u64* M, size_t M_ls, size_t M_rs, uint32_t rows, uint32_t k, uint32_t L, uint32_t ell, u64* out,
                            size_t out_plane, bool packed, cudaStream_t st) {
  if (rows == 0) return;
  imma_slot_major_kernel<<<dim3(rows, L), 256, 0, st>>>(M, M_ls, M_rs, k, ell, out, out_plane, packed ? 1 : 0);
}

}  // namespace pvw
