// imma.cu -- the modular matrix product on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
//   acc[row][d] = sum_{j<k} M[row][j] * V[d][j]  (mod q)      per plane = (RNS limb, NTT slot); M, V < 2^62
//
// Batched over dealers the product is a GEMM per plane and, on the CUDA cores, bound by the integer pipe (mac.cu: three
// IMAD.WIDE per 62-bit multiply-accumulate, 2.8e12 MAC/s ceiling).  Exact integer GEMMs are what the INT8 tensor cores do:
// with m_s, v_t the little-endian bytes of the operands,
//     M * V = sum_{u=0..14} 2^(8u) * S_u ,      S_u = sum_{s+t=u} sum_j m_s(j) v_t(j)   (< 8 k 255^2 < 2^31 for k <= 4096).
// Both operands are stored as byte planes.  For one tile (128 rows x DT dealers) the B operand is the tile of V's byte
// planes, rows (t, d) -- 8*DT rows, one 128-byte K-chunk at a time in shared memory -- and for every byte plane s of M one
// chain of MMAs  D[:, DT*s .. DT*s + 8*DT) += M_s * B^T  runs on a WINDOW of the accumulator that starts DT*s columns in:
// row (t, d) of B lands in column DT*(s+t) + d, so the 64 byte products of every output add up, in place, to its 15
// diagonal sums S_u at columns DT*u + d.  64 int8 multiply-accumulates per 62-bit one, no expanded operand, no zero
// padding in the GEMM.  The epilogue thread that owns a TMEM lane (= a row) reads the 15 sums of each dealer,
// recombines them into one 160-bit integer and reduces it: no cross-lane traffic.
// (First version, kept in git history: V expanded into 15 diagonal rows, M rows used as they lie in memory -- 120 int8
//  MACs per 62-bit one and 15x the V bytes; 8.4e12 MAC/s.  This one: 2.05e13 MAC/s on the c2 product of the bench.)
//
// Kernel shape: persistent, one CTA per SM, warp specialised --
//   warp 0: TMA producer.  RES (two whole B tiles fit, k <= 256): the B tile once per output tile into one of two slots, then
//           the (plane s, K-chunk) tiles of M, plane-major, through a ring of 16 KB stages; otherwise one 128-byte K-chunk of
//           B at a time (two chunk slots) and the loop runs chunk-major, so shared memory does not depend on k.
//   warp 1: TMEM allocation + MMA issue by one lane (elect.sync).  This thread's scalar code paces the whole kernel: ring
//           position / phase are counters and descriptors are base + offset -- with two integer divisions per stage and
//           descriptors rebuilt from addresses every MMA cost ~195 cycles whatever its shape.
//   warps 2-9: epilogue, two warps per TMEM lane group, half the dealers each: phase 1 turns the 15 sums of each dealer into
//           five 32-bit words and hands TMEM back, phase 2 (reduction, stores) runs under the next tile's MMAs.  Measured
//           alternative (not kept): reading diagonal u as soon as plane u has completed and freeing its columns at once --
//           slower with both the old and the lean issue loop (2.16 ms against 1.98 ms per c2 launch).
// A window that starts DT*s columns in touches DT columns no earlier MMA has written: on the first K step of plane s >= 1 the
// MMA is issued in two parts, N = 7*DT accumulating and N = DT (the t = 7 rows of B) overwriting, so TMEM never needs clearing.
// Outputs should be written slot-major (O_rs = 1): the lanes of a warp are consecutive rows, so that is the unit-stride form.
#include <cuda.h>

#include <algorithm>

#include "imma.cuh"

namespace pvw {

namespace {

constexpr uint32_t RT = 128, KC = 128, A_STAGE = RT * KC;
constexpr uint32_t EPI_WARPS = 8, THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t SMEM_LIMIT = 227 * 1024;
constexpr uint32_t MAX_STAGES = 10;
constexpr uint32_t BAR_BYTES = 256;

#define IMMA_DEV __device__ __forceinline__

IMMA_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
IMMA_DEV void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
IMMA_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
IMMA_DEV void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
IMMA_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();  // never hang the device: a lost barrier becomes a launch error
  } while (!ok);
}
IMMA_DEV void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
IMMA_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
IMMA_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
IMMA_DEV void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32, M = 128, N from the instruction descriptor, K = 32
IMMA_DEV void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// one lane of a converged warp, chosen by the hardware: the form under which ptxas keeps the single-thread tcgen05 issue
// code on the uniform datapath (a plain `lane == 0` branch wraps every UTCIMMA in an ELECT / BRA.U.ANY loop)
IMMA_DEV bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
IMMA_DEV void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (SBO), version 1 (sm_100)
IMMA_DEV uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (kind::i8): D = s32 (bits 4-5 = 2), A and B unsigned 8-bit (bits 7-9, 10-12 = 0), both K-major
// (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28
__host__ __device__ constexpr uint32_t idesc_n(uint32_t n) { return (2u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

// W = sum_u s[u] * 2^(8u) as five 32-bit words (s[u] < 2^32; the true value is < k * 2^124, so the fifth word stays small)
IMMA_DEV void combine160(const uint32_t (&s)[IMMA_DIAGS], uint32_t (&w)[5]) {
  // diagonals u = r mod 4 are word aligned among themselves: X_r = (s[r], s[4+r], s[8+r], s[12+r]) as a 128-bit integer, and the
  // value is X_0 + (X_1 << 8) + (X_2 << 16) + (X_3 << 24)
  u64 acc = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const uint32_t c0 = i < 4 ? s[4 * i] : 0u, c1 = i < 4 ? s[4 * i + 1] : 0u, c2 = i < 4 ? s[4 * i + 2] : 0u, c3 = i < 3 ? s[4 * i + 3] : 0u;
    const uint32_t p1 = i > 0 ? s[4 * i - 3] : 0u, p2 = i > 0 ? s[4 * i - 2] : 0u, p3 = (i > 0 && i < 4) ? s[4 * i - 1] : 0u;
    acc += (u64)c0 + __funnelshift_l(p1, c1, 8) + __funnelshift_l(p2, c2, 16) + __funnelshift_l(p3, c3, 24);
    w[i] = (uint32_t)acc;
    acc >>= 32;
  }
}
IMMA_DEV u64 reduce160(const uint32_t (&w)[5], const LimbConst& lc) {
  const u64 lo = ((u64)w[1] << 32) | w[0], hi = ((u64)w[3] << 32) | w[2];
  const u64 h = reduce128(reduce64((u64)w[4], lc), hi, lc);
  return reduce128(h, lo, lc);
}
// The same for q >= 2^61 and W < 2^136 (k <= 4096), one Barrett step instead of two and a half: the bits above 2^124 (at most 12)
// are folded down with 2^124 mod q, which leaves t < 2^124 + 2^74, and t / q < 2^64 is what reduce128's quotient estimate needs.
IMMA_DEV u64 reduce160_q62(const uint32_t (&w)[5], const LimbConst& lc) {
  const uint32_t h = (w[3] >> 28) | (w[4] << 4);
  const u64 lo = ((u64)w[1] << 32) | w[0], hi = ((u64)(w[3] & 0x0fffffffu) << 32) | w[2];
  const u64 p0 = (u64)h * (uint32_t)lc.c124, p1 = (u64)h * (uint32_t)(lc.c124 >> 32);      // h * c124 = p0 + (p1 << 32) < 2^74
  const u64 a = lo + p0, b = a + (p1 << 32);
  const u64 t_hi = hi + (p1 >> 32) + (a < lo) + (b < a);
  return reduce128(t_hi, b, lc);
}
IMMA_DEV void tc_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

// DT dealers per tile (32: 15 * 32 = 480 of the 512 TMEM columns)
// EW epilogue warps (8 or 16): EW / 4 warps share a TMEM lane group and split the DT dealers
// Q62: every modulus of the launch is >= 2^61 (decided on the host): the one-step reduction.  A template parameter, not a per-tile
// branch: with both reductions in the unrolled store loop the epilogue code grew by a third and the general path ran 17 % slower
// (instruction fetch), which cancelled what the shorter reduction gains.
template <uint32_t DT, bool RES, uint32_t EW, bool Q62>
__global__ void __launch_bounds__(64 + 32 * EW, 1) imma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                               const ImmaArgs g, const uint32_t nstages) {
  constexpr uint32_t resident = RES ? 1u : 0u;
  // resident != 0: a B slot holds the whole B tile (all K-chunks) and the loop runs plane-major (s outer, K-chunk inner): one B
  // load per tile, the next tile's B arrives during this tile's MMAs -- used when two such slots fit (k <= 256 at DT = 32).
  // resident == 0: a B slot holds one K-chunk and the loop runs chunk-major; shared memory does not depend on k.
  constexpr uint32_t NB = 8 * DT;                 // rows of the B tile = MMA columns per window
  constexpr uint32_t B_CHUNK = NB * KC;           // one K-chunk of the B tile (multiple of 1024 bytes)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // 128-byte swizzle wants 1024-byte aligned tiles
  const uint32_t kp = imma_kp(g.k), nkc = (kp + KC - 1) / KC;
  const uint32_t bunit = resident ? nkc * B_CHUNK : B_CHUNK;            // bytes of one B slot; two slots, then the M ring
  const uint32_t b_base = base, a_base = base + 2 * bunit, bar0 = a_base + nstages * A_STAGE;
  uint8_t* gen = smem_raw + (bar0 - smem_u32(smem_raw));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 8 * (2 * MAX_STAGES + 6));
  auto a_full = [&](uint32_t s) { return bar0 + 8 * s; };
  auto a_empty = [&](uint32_t s) { return bar0 + 8 * (MAX_STAGES + s); };
  auto b_full = [&](uint32_t b) { return bar0 + 8 * (2 * MAX_STAGES + b); };
  auto b_empty = [&](uint32_t b) { return bar0 + 8 * (2 * MAX_STAGES + 2 + b); };
  const uint32_t tmem_full = bar0 + 8 * (2 * MAX_STAGES + 4), tmem_empty = tmem_full + 8;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent: CTA b takes tiles b, b + gridDim.x, ...; row tiles vary fastest, then dealer tiles, then planes, so the CTAs
  // running at any moment share one plane's operands in L2
  const uint32_t n_rt = (g.rows + RT - 1) / RT, n_dt = (g.D + DT - 1) / DT, total = n_rt * n_dt * g.L * g.ell;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < nstages; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (uint32_t b = 0; b < 2; b++) { mbar_init(b_full(b), 1); mbar_init(b_empty(b), 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 32 * EW);                                      // every epilogue thread arrives
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
      uint32_t st = 0, ph = 1, bcnt = 0, loaded = 0;                      // ph: parity to wait for on the empty barriers
      for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const uint32_t rt = tile % n_rt, dt = (tile / n_rt) % n_dt, plane = tile / (n_rt * n_dt);
        auto load_b = [&](uint32_t kc0, uint32_t kc1) {                  // K-chunks [kc0, kc1) of the B tile into the next B slot
          const uint32_t bb = bcnt & 1, bit = bcnt >> 1;
          mbar_wait(b_empty(bb), (bit & 1) ^ 1);                         // the MMAs that read this slot two fills ago are done
          if (g.mode >= 4 && bcnt >= 2) mbar_arrive(b_full(bb));         // timing probe: operands stay as loaded the first time
          else {
            mbar_expect_tx(b_full(bb), bunit);
            for (uint32_t kc = kc0; kc < kc1; kc++)
              tma_load_4d(b_base + bb * bunit + (RES ? kc * B_CHUNK : 0u), &tmB, (int)(kc * KC), (int)(g.d_first + dt * DT), 0, (int)plane, b_full(bb));
          }
          bcnt++;
        };
        auto load_a = [&](uint32_t s, uint32_t kc) {
          mbar_wait(a_empty(st), ph);                                    // first pass over the ring: passes at once
          if (g.mode >= 4 && loaded >= nstages) mbar_arrive(a_full(st));
          else {
            mbar_expect_tx(a_full(st), A_STAGE);
            tma_load_4d(a_base + st * A_STAGE, &tmA, (int)(kc * KC), (int)s, (int)(rt * RT), (int)plane, a_full(st));
            loaded++;
          }
          if (++st == nstages) { st = 0; ph ^= 1; }
        };
        if (RES) {
          load_b(0, nkc);
          for (uint32_t s = 0; s < 8; s++)
            for (uint32_t kc = 0; kc < nkc; kc++) load_a(s, kc);
        } else {
          for (uint32_t kc = 0; kc < nkc; kc++) {
            load_b(kc, kc + 1);
            for (uint32_t s = 0; s < 8; s++) load_a(s, kc);
          }
        }
      }
    }
  } else if (warp == 1) {
    uint32_t st = 0, ph = 0, bcnt = 0, i = 0;
    const uint32_t a_lo = (a_base & 0x3FFFFu) >> 4, b_lo = (b_base & 0x3FFFFu) >> 4;
    for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x, i++) {
      mbar_wait(tmem_empty, (i & 1) ^ 1);                                // the epilogue has read the previous tile out of TMEM
      tc_fence_after();
      // The issuing lane is a single thread running scalar code between the MMAs: measured, the MMA stream is paced by THIS
      // loop, not by the tensor pipe or the loads (a version with two integer divisions per stage ran 25 % slower).  So: ring
      // position and phase are counters, descriptors are base + offset in their 14-bit address field, nothing is divided.
      constexpr uint64_t DESC_HI = ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
      uint32_t bb = 0;
      auto step = [&](uint32_t s, uint32_t kc) {
        mbar_wait(a_full(st), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = DESC_HI | (uint64_t)(a_lo + st * (A_STAGE >> 4));
          const uint64_t bdesc = DESC_HI | (uint64_t)(b_lo + bb * (bunit >> 4) + (RES ? kc * (B_CHUNK >> 4) : 0u));
          const uint32_t d_addr = tmem_base + DT * s;
          if (kc != 0) {
#pragma unroll
            for (uint32_t k4 = 0; k4 < KC / 32; k4++) tc_mma_i8(d_addr, adesc + 2 * k4, bdesc + 2 * k4, idesc_n(NB), 1);
          } else {
            if (s == 0) tc_mma_i8(d_addr, adesc, bdesc, idesc_n(NB), 0);
            else {  // columns DT*(s+7) .. +DT are new to this window: their first product overwrites
              tc_mma_i8(d_addr, adesc, bdesc, idesc_n(7 * DT), 1);
              tc_mma_i8(d_addr + 7 * DT, adesc, bdesc + (7 * DT * KC >> 4), idesc_n(DT), 0);
            }
#pragma unroll
            for (uint32_t k4 = 1; k4 < KC / 32; k4++) tc_mma_i8(d_addr, adesc + 2 * k4, bdesc + 2 * k4, idesc_n(NB), 1);
          }
          tc_commit(a_empty(st));                                        // arrives when the MMAs above have read the stage
        }
        __syncwarp();
        if (++st == nstages) { st = 0; ph ^= 1; }
      };
      auto next_b = [&]() {
        bb = bcnt & 1;
        mbar_wait(b_full(bb), (bcnt >> 1) & 1);
        bcnt++;
      };
      if (RES) {
        next_b();
        for (uint32_t s = 0; s < 8; s++)
          for (uint32_t kc = 0; kc < nkc; kc++) step(s, kc);
        if (lane == 0) tc_commit(b_empty(bb));
      } else {
        for (uint32_t kc = 0; kc < nkc; kc++) {
          next_b();
          for (uint32_t s = 0; s < 8; s++) step(s, kc);
          if (lane == 0) tc_commit(b_empty(bb));
        }
      }
      if (lane == 0) tc_commit(tmem_full);
      __syncwarp();
    }
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w % 4) .. +31; thread = one row; the EW / 4 warps of a lane group split the dealers
    const uint32_t lg = warp & 3, half = (warp - 2) >> 2;
    uint32_t i = 0;
    for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x, i++) {
      const uint32_t rt = tile % n_rt, dt = (tile / n_rt) % n_dt, plane = tile / (n_rt * n_dt);
      const uint32_t row = rt * RT + lg * 32 + lane, d0 = dt * DT;
      const uint32_t limb = plane / g.ell, c = plane - limb * g.ell;
      const LimbConst lc = g.lc[limb];
      const bool row_ok = row < g.rows;
      const size_t o_row = (size_t)limb * g.O_ls + (size_t)row * g.O_rs + (size_t)c * g.O_cs;
      const uint32_t srow = (g.mode == 1 && row_ok) ? (g.S_rowmap ? g.S_rowmap[row] : row) : 0;
      mbar_wait(tmem_full, i & 1);
      tc_fence_after();
      if (g.mode >= 3) {                                                 // timing probes: MMA / load pipeline without an epilogue
        tc_fence_before();
        mbar_arrive(tmem_empty);
        continue;
      }
      // phase 1 (TMEM is busy): this thread's DT/2 dealers, four at a time -- 15 diagonal sums each -> one 160-bit integer
      constexpr uint32_t ND = DT / (EW / 4);
      uint32_t W[ND][5];
#pragma unroll
      for (uint32_t q4 = 0; q4 < ND / 4; q4++) {
        uint32_t v[IMMA_DIAGS][4];
#pragma unroll
        for (uint32_t u = 0; u < IMMA_DIAGS; u++) tc_ld4(tmem_base + ((lg * 32) << 16) + DT * u + half * ND + q4 * 4, v[u]);
        tc_ld_wait();
#pragma unroll
        for (uint32_t dd = 0; dd < 4; dd++) {
          uint32_t s[IMMA_DIAGS];
#pragma unroll
          for (uint32_t u = 0; u < IMMA_DIAGS; u++) s[u] = v[u][dd];
          combine160(s, W[q4 * 4 + dd]);
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);                                           // TMEM goes back to the MMA warp: the next tile starts
      // phase 2 (overlaps the next tile's MMAs): reduce and store
#pragma unroll
      for (uint32_t dd = 0; dd < ND; dd++) {
        const uint32_t d = d0 + half * ND + dd;
        if (row_ok && d < g.D) {
          // (mode -1, timing probe: the read-out and the stores without the reduction)
          u64 r = g.mode == -1 ? ((u64)(W[dd][4] ^ W[dd][3] ^ W[dd][2]) << 32 | (W[dd][1] ^ W[dd][0])) : Q62 ? reduce160_q62(W[dd], lc) : reduce160(W[dd], lc);
          u64* o = g.O + (size_t)d * g.O_ds + o_row;
          if (g.mode == 0) r = addmod(r, *o, lc.q);
          else if (g.mode == 1) {
            const uint32_t sd = g.V_dmap ? g.V_dmap[d] : d;
            r = submod(r, g.S[(size_t)sd * g.S_ds + (size_t)limb * g.S_ls + (size_t)srow * g.ell + c], lc.q);
          }
          *o = g.O_packed ? pack_halves(r) : r;
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

// ---------------------------------------------------------------------------------------------------------------------
// Two-SM form (tcgen05.mma.cta_group::2) -- MEASURED ALTERNATIVE, off by default (option "imma_pair" / PVW_IMMA_PAIR=1).
// The MMA stream of the single-CTA kernel runs at ~195-215 cycles per instruction whatever N (128 or 256) and whether or
// not anything is loaded (probe modes 3-5), i.e. 59 % of the nominal int8 rate.  If the shared-memory operand fetch (4 KB
// of A + 8 KB of B per instruction) were the cause, a CTA pair sharing one M = 256 MMA would help: each CTA stages its own
// 128 rows of M and HALF of the B tile (byte planes t = 4*rank .. 4*rank+3 of the DT dealers), 8 KB per SM and instruction.
// It does not: the pair kernel is bit-exact and exactly as fast (2.37 ms per c2 launch against 2.36 ms), so the stream is
// at what this part sustains under tensor load -- the same 60 % of nominal that MEASURED_PEAKS.json records for cuBLAS bf16
// (1 355 of 2 250 TFLOP/s sustained).  The leader CTA issues the MMAs for both;
// TMA loads of both CTAs complete on the leader's full barriers, tcgen05.commit multicasts the empty / done barriers to both.
// The N = 7*DT + DT split of the first K step does not exist here (B is split evenly between the CTAs), so the epilogue
// clears the upper columns (diagonals 8..14) with tcgen05.st after reading them.
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same offset in CTA rank 0 of the pair

IMMA_DEV uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
IMMA_DEV void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
IMMA_DEV void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
IMMA_DEV void tc_mma_i8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
IMMA_DEV void tc_commit_2sm(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
IMMA_DEV void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
IMMA_DEV void tc_st8_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
IMMA_DEV void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// M = 256 (two CTAs x 128 lanes): M >> 4 = 16 at bits 24-28
__host__ __device__ constexpr uint32_t idesc2_n(uint32_t n) { return (2u << 4) | ((n >> 3) << 17) | ((256u >> 4) << 24); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
imma_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ImmaArgs g, const uint32_t nstages) {
  constexpr uint32_t DT = 32, NB = 8 * DT, B_HALF = (NB / 2) * KC;      // this CTA's half of one K-chunk of the B tile: 16 KB
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t kp = imma_kp(g.k), nkc = (kp + KC - 1) / KC;
  const uint32_t b_base = base, a_base = base + 2 * B_HALF, bar0 = a_base + nstages * A_STAGE;
  uint8_t* gen = smem_raw + (bar0 - smem_u32(smem_raw));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 8 * (2 * MAX_STAGES + 6));
  auto a_full = [&](uint32_t s) { return bar0 + 8 * s; };
  auto a_empty = [&](uint32_t s) { return bar0 + 8 * (MAX_STAGES + s); };
  auto b_full = [&](uint32_t b) { return bar0 + 8 * (2 * MAX_STAGES + b); };
  auto b_empty = [&](uint32_t b) { return bar0 + 8 * (2 * MAX_STAGES + 2 + b); };
  const uint32_t tmem_full = bar0 + 8 * (2 * MAX_STAGES + 4), tmem_empty = tmem_full + 8;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rank = cluster_rank();
  const bool leader = rank == 0;
  // pair tile = 256 rows x DT dealers x plane; the pair b / 2 walks pair tiles b/2, b/2 + gridDim.x/2, ...
  const uint32_t n_rt = (g.rows + 2 * RT - 1) / (2 * RT), n_dt = (g.D + DT - 1) / DT, total = n_rt * n_dt * g.L * g.ell;
  const uint32_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < nstages; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (uint32_t b = 0; b < 2; b++) { mbar_init(b_full(b), 1); mbar_init(b_empty(b), 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 2 * 32 * EPI_WARPS);                           // the epilogue threads of BOTH CTAs arrive on the leader's
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                                    // barrier inits and TMEM of both CTAs are visible
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 1, bcnt = 0;                                   // ph: parity to wait for on the empty barriers
      for (uint32_t tile = pair; tile < total; tile += npairs) {
        const uint32_t rt = tile % n_rt, dt = (tile / n_rt) % n_dt, plane = tile / (n_rt * n_dt);
        for (uint32_t kc = 0; kc < nkc; kc++, bcnt++) {
          const uint32_t bb = bcnt & 1, bit = bcnt >> 1;
          mbar_wait(b_empty(bb), (bit & 1) ^ 1);
          if (leader) mbar_expect_tx(b_full(bb), 2 * B_HALF);            // both halves complete on the leader's barrier
          tma_load_4d_2sm(b_base + bb * B_HALF, &tmB, (int)(kc * KC), (int)(g.d_first + dt * DT), (int)(4 * rank), (int)plane, b_full(bb) & PEER_MASK);
          for (uint32_t s = 0; s < 8; s++) {
            mbar_wait(a_empty(st), ph);
            if (leader) mbar_expect_tx(a_full(st), 2 * A_STAGE);
            tma_load_4d_2sm(a_base + st * A_STAGE, &tmA, (int)(kc * KC), (int)s, (int)(rt * 2 * RT + rank * RT), (int)plane, a_full(st) & PEER_MASK);
            if (++st == nstages) { st = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // the same lean issue loop as the one-SM kernel: counters for ring position and phase, descriptors as base + offset, the
      // issuing lane chosen by elect.sync (the first version of this kernel rebuilt descriptors and divided per stage, and ran at
      // the pace of that scalar code, which hid what the pair buys: each SM fetches 4 + 4 KB of operands per MMA instead of 4 + 8)
      constexpr uint64_t DESC_HI = ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t a_lo = (a_base & 0x3FFFFu) >> 4, b_lo = (b_base & 0x3FFFFu) >> 4;
      uint32_t st = 0, ph = 0, bcnt = 0, i = 0;
      for (uint32_t tile = pair; tile < total; tile += npairs, i++) {
        mbar_wait(tmem_empty, i & 1);                                    // both epilogues have read and cleared the previous tile (phase 0: the initial clearing)
        tc_fence_after();
        for (uint32_t kc = 0; kc < nkc; kc++, bcnt++) {
          const uint32_t bb = bcnt & 1, bit = bcnt >> 1;
          mbar_wait(b_full(bb), bit & 1);
          for (uint32_t s = 0; s < 8; s++) {
            mbar_wait(a_full(st), ph);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t adesc = DESC_HI | (uint64_t)(a_lo + st * (A_STAGE >> 4));
              const uint64_t bdesc = DESC_HI | (uint64_t)(b_lo + bb * (B_HALF >> 4));
              const uint32_t d_addr = tmem_base + DT * s;
              if ((kc | s) == 0) {
                tc_mma_i8_2sm(d_addr, adesc, bdesc, idesc2_n(NB), 0);
#pragma unroll
                for (uint32_t k4 = 1; k4 < KC / 32; k4++) tc_mma_i8_2sm(d_addr, adesc + 2 * k4, bdesc + 2 * k4, idesc2_n(NB), 1);
              } else {
#pragma unroll
                for (uint32_t k4 = 0; k4 < KC / 32; k4++) tc_mma_i8_2sm(d_addr, adesc + 2 * k4, bdesc + 2 * k4, idesc2_n(NB), 1);
              }
              tc_commit_2sm(a_empty(st));
              if (s == 7) tc_commit_2sm(b_empty(bb));
            }
            __syncwarp();
            if (++st == nstages) { st = 0; ph ^= 1; }
          }
        }
        if (elect_one()) tc_commit_2sm(tmem_full);
        __syncwarp();
      }
    }
  } else {
    const uint32_t lg = warp & 3, half = (warp - 2) >> 2;
    // diagonals 8..14 start from zero in every tile (the first MMA of a tile only overwrites the window of plane 0)
    for (uint32_t col = 8 * DT + half * 8; col < IMMA_DIAGS * DT; col += 16) tc_st8_zero(tmem_base + ((lg * 32) << 16) + col);
    tc_st_wait();
    tc_fence_before();
    mbar_arrive_cluster(tmem_empty & PEER_MASK);
    uint32_t i = 0;
    for (uint32_t tile = pair; tile < total; tile += npairs, i++) {
      const uint32_t rt = tile % n_rt, dt = (tile / n_rt) % n_dt, plane = tile / (n_rt * n_dt);
      const uint32_t row = rt * 2 * RT + rank * RT + lg * 32 + lane, d0 = dt * DT;
      const uint32_t limb = plane / g.ell, c = plane - limb * g.ell;
      const LimbConst lc = g.lc[limb];
      const bool row_ok = row < g.rows;
      const size_t o_row = (size_t)limb * g.O_ls + (size_t)row * g.O_rs + (size_t)c * g.O_cs;
      const uint32_t srow = (g.mode == 1 && row_ok) ? (g.S_rowmap ? g.S_rowmap[row] : row) : 0;
      mbar_wait(tmem_full, i & 1);
      tc_fence_after();
      constexpr uint32_t ND = DT / 2;
      uint32_t W[ND][5];
#pragma unroll
      for (uint32_t q4 = 0; q4 < ND / 4; q4++) {
        uint32_t v[IMMA_DIAGS][4];
#pragma unroll
        for (uint32_t u = 0; u < IMMA_DIAGS; u++) tc_ld4(tmem_base + ((lg * 32) << 16) + DT * u + half * ND + q4 * 4, v[u]);
        tc_ld_wait();
#pragma unroll
        for (uint32_t dd = 0; dd < 4; dd++) {
          uint32_t s[IMMA_DIAGS];
#pragma unroll
          for (uint32_t u = 0; u < IMMA_DIAGS; u++) s[u] = v[u][dd];
          combine160(s, W[q4 * 4 + dd]);
        }
      }
      // clear this thread's part of diagonals 8..14 for the next tile, then hand TMEM back
#pragma unroll
      for (uint32_t u = 8; u < IMMA_DIAGS; u++) {
        tc_st8_zero(tmem_base + ((lg * 32) << 16) + DT * u + half * ND);
        tc_st8_zero(tmem_base + ((lg * 32) << 16) + DT * u + half * ND + 8);
      }
      tc_st_wait();
      tc_fence_before();
      mbar_arrive_cluster(tmem_empty & PEER_MASK);
      if (g.mode >= 3) continue;
#pragma unroll
      for (uint32_t dd = 0; dd < ND; dd++) {
        const uint32_t d = d0 + half * ND + dd;
        if (row_ok && d < g.D) {
          u64 r = reduce160(W[dd], lc);
          u64* o = g.O + (size_t)d * g.O_ds + o_row;
          if (g.mode == 0) r = addmod(r, *o, lc.q);
          else if (g.mode == 1) {
            const uint32_t sd = g.V_dmap ? g.V_dmap[d] : d;
            r = submod(r, g.S[(size_t)sd * g.S_ds + (size_t)limb * g.S_ls + (size_t)srow * g.ell + c], lc.q);
          }
          *o = g.O_packed ? pack_halves(r) : r;
        }
      }
    }
    tc_fence_before();
  }
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// Byte planes of a limb-major operand.  One CTA per (operand row, limb, chunk of JC polynomials): the chunk is read with
// unit stride into shared memory (rows padded by one word: the transposed reads below are conflict free), then each thread
// takes one slot c of FOUR consecutive polynomials j and writes the eight byte planes as 4-byte words -- a warp stores 128
// contiguous bytes per instruction (the first version stored single bytes: 8x the store instructions, 32-byte bursts).
//   src  = M + rowmap(row) * M_rs + limb * M_ls + j * ell + c            (canonical residues, or packed halves)
//   dst  = Mb + (limb * ell + c) * Mb_plane + row * dst_rs + s * dst_bs + j          s = byte index
// both sides (A, B, s_hat; r_hat, c1): dst_rs = 8 * kp, dst_bs = kp.  (The dealer side first kept byte plane t of all dealers
// together, rows * kp apart: 64 stores per polynomial 512 KB apart at 2048 dealers ran at a third of the matrix side's rate.)
IMMA_DEV void transpose_4x8(const u64 (&v)[4], uint32_t (&w)[8]) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t a0 = (uint32_t)(v[0] >> (32 * h)), a1 = (uint32_t)(v[1] >> (32 * h)), a2 = (uint32_t)(v[2] >> (32 * h)), a3 = (uint32_t)(v[3] >> (32 * h));
    const uint32_t t01 = __byte_perm(a0, a1, 0x5140), t23 = __byte_perm(a2, a3, 0x5140);   // bytes 0, 1 of the pair, interleaved
    const uint32_t u01 = __byte_perm(a0, a1, 0x7362), u23 = __byte_perm(a2, a3, 0x7362);   // bytes 2, 3
    w[4 * h + 0] = __byte_perm(t01, t23, 0x5410);
    w[4 * h + 1] = __byte_perm(t01, t23, 0x7632);
    w[4 * h + 2] = __byte_perm(u01, u23, 0x5410);
    w[4 * h + 3] = __byte_perm(u01, u23, 0x7632);
  }
}

__global__ void __launch_bounds__(256) imma_planes_kernel(const u64* __restrict__ M, size_t M_ls, size_t M_rs, uint32_t k, uint32_t ell, uint32_t jc,
                                                          uint8_t* __restrict__ Mb, size_t Mb_plane, size_t dst_rs, size_t dst_bs, int packed,
                                                          const uint32_t* __restrict__ rowmap, uint32_t nchunks) {
  extern __shared__ u64 s_v[];                                            // [jc][ell + 1]
  const uint32_t row = blockIdx.x, limb = blockIdx.y / nchunks, j0 = (blockIdx.y - limb * nchunks) * jc;
  const uint32_t jn = min(jc, k - j0), pitch = ell + 1;
  const u64* src = M + (size_t)(rowmap ? rowmap[row] : row) * M_rs + (size_t)limb * M_ls + (size_t)j0 * ell;
  for (uint32_t t = threadIdx.x; t < jn * ell; t += blockDim.x) {
    u64 v = src[t];
    if (packed) v = unpack_halves(v);
    const uint32_t j = t / ell, c = t - j * ell;
    s_v[j * pitch + c] = v;
  }
  __syncthreads();
  const uint32_t groups = (jn + 3) / 4;
  for (uint32_t t = threadIdx.x; t < groups * ell; t += blockDim.x) {
    const uint32_t c = t / groups, jg = t - c * groups;
    u64 v[4];
#pragma unroll
    for (uint32_t i = 0; i < 4; i++) v[i] = (4 * jg + i < jn) ? s_v[(4 * jg + i) * pitch + c] : 0ull;   // the padding of a plane row is zero
    uint32_t w[8];
    transpose_4x8(v, w);
    uint8_t* out = Mb + (size_t)(limb * ell + c) * Mb_plane + (size_t)row * dst_rs + j0 + 4 * jg;
#pragma unroll
    for (uint32_t b = 0; b < 8; b++) *reinterpret_cast<uint32_t*>(out + (size_t)b * dst_bs) = w[b];
  }
}

// the inverse: byte planes -> limb-major operand (used when the u64 copy of B was released and a single call / download needs it)
__global__ void __launch_bounds__(256) imma_unplanes_kernel(u64* __restrict__ M, size_t M_ls, size_t M_rs, uint32_t k, uint32_t ell, uint32_t jc,
                                                            const uint8_t* __restrict__ Mb, size_t Mb_plane, size_t src_rs, size_t src_bs, int packed,
                                                            uint32_t nchunks) {
  extern __shared__ u64 s_v[];                                            // [jc][ell + 1]
  const uint32_t row = blockIdx.x, limb = blockIdx.y / nchunks, j0 = (blockIdx.y - limb * nchunks) * jc;
  const uint32_t jn = min(jc, k - j0), pitch = ell + 1, groups = (jn + 3) / 4;
  for (uint32_t t = threadIdx.x; t < groups * ell; t += blockDim.x) {
    const uint32_t c = t / groups, jg = t - c * groups;
    const uint8_t* in = Mb + (size_t)(limb * ell + c) * Mb_plane + (size_t)row * src_rs + j0 + 4 * jg;
    uint32_t w[8];
#pragma unroll
    for (uint32_t b = 0; b < 8; b++) w[b] = *reinterpret_cast<const uint32_t*>(in + (size_t)b * src_bs);
#pragma unroll
    for (uint32_t i = 0; i < 4; i++) {
      u64 v = 0;
#pragma unroll
      for (uint32_t b = 0; b < 8; b++) v |= (u64)((w[b] >> (8 * i)) & 0xffu) << (8 * b);
      if (4 * jg + i < jn) s_v[(4 * jg + i) * pitch + c] = v;
    }
  }
  __syncthreads();
  u64* dst = M + (size_t)row * M_rs + (size_t)limb * M_ls + (size_t)j0 * ell;
  for (uint32_t t = threadIdx.x; t < jn * ell; t += blockDim.x) {
    const uint32_t j = t / ell, c = t - j * ell;
    const u64 v = s_v[j * pitch + c];
    dst[t] = packed ? pack_halves(v) : v;
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
// 4D u8 tensor, innermost dimension contiguous, 128-byte swizzle, out-of-range elements read as zero
bool make_map(CUtensorMap* tm, const void* base, const cuuint64_t (&dims)[4], const cuuint64_t (&strides)[3], const cuuint32_t (&box)[4]) {
  encode_fn encode = tensor_map_encoder();
  if (!encode || ((uintptr_t)base & 15)) return false;
  for (cuuint64_t s : strides) if (s & 15) return false;
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <uint32_t DT>
bool launch_dt(const ImmaArgs& a, cudaStream_t st) {
  const uint32_t kp = imma_kp(a.k), nkc = (kp + KC - 1) / KC;
  const uint64_t planes = (uint64_t)a.L * a.ell;
  CUtensorMap tmA, tmB;
  // M byte planes: [plane][row][s][kp]; box = 128 bytes of one plane s of 128 rows
  if (!make_map(&tmA, a.Mb, {kp, 8, a.rows, planes}, {kp, 8ull * kp, a.Mb_plane}, {KC, 1, RT, 1})) return false;
  // V byte planes: [plane][d][t][kp] (the same layout as the matrix side: a dealer's eight planes are 8 * kp contiguous bytes);
  // box = 128 bytes of DT dealers of all 8 planes t, t outermost -> rows (t, d) of the B tile in shared memory
  if (!make_map(&tmB, a.Vb, {kp, a.Vb_D, 8, planes}, {a.Vb_dstride ? (cuuint64_t)a.Vb_dstride : 8ull * kp, kp, a.Vb_plane}, {KC, DT, 8, 1})) return false;
  // two whole B tiles resident when at least four ring stages still fit (k <= 256 at DT = 32), else two K-chunk slots
  const uint32_t b_tile = nkc * 8 * DT * KC;
  const uint32_t resident = (2 * b_tile + 4 * A_STAGE + 1024 + BAR_BYTES <= SMEM_LIMIT) ? 1u : 0u;
  const uint32_t b_bytes = resident ? 2 * b_tile : 2 * 8 * DT * KC;
  const uint32_t cap = a.stages >= 2 ? std::min<uint32_t>((uint32_t)a.stages, MAX_STAGES) : MAX_STAGES;
  const uint32_t nstages = std::max(2u, std::min<uint32_t>(cap, (SMEM_LIMIT - 1024 - BAR_BYTES - b_bytes) / A_STAGE));
  const uint32_t smem = b_bytes + nstages * A_STAGE + 1024 + BAR_BYTES;
  // 8 epilogue warps unless asked otherwise.  Measured (c2-sized product, one box): 8 warps + the general reduction 1.989 ms, 16 warps
  // 1.809 (the reductions of a tile are latency bound with two warps per scheduler); with the one-step reduction 1.732 / 1.759 -- the
  // read-out itself is slower with 16 warps (1.664 against 1.564 ms without any reduction), so 8 stays the default
  const bool wide = a.epi_warps == 16;
  auto pick = [&](auto q62) {
    constexpr bool Q = decltype(q62)::value;
    return wide ? (resident ? imma_gemm_kernel<DT, true, 16, Q> : imma_gemm_kernel<DT, false, 16, Q>)
                : (resident ? imma_gemm_kernel<DT, true, 8, Q> : imma_gemm_kernel<DT, false, 8, Q>);
  };
  auto kern = a.fast_reduce ? pick(std::true_type{}) : pick(std::false_type{});
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return false;  // per device
  // the whole 228 KB as shared memory whatever this launch asks for: with a shorter ring (ImmaArgs::stages) the rest stays free for
  // the CTAs of kernels on other streams (the driver would otherwise pick the smallest carve-out that fits this kernel alone)
  if (cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) return false;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
  const uint64_t tiles = (uint64_t)((a.rows + RT - 1) / RT) * ((a.D + DT - 1) / DT) * planes;
  if (tiles >= (1ull << 32)) return false;
  kern<<<(unsigned)std::min<uint64_t>(tiles, (uint64_t)std::max(sms, 1)), 64 + 32 * (wide ? 16 : 8), smem, st>>>(tmA, tmB, a, nstages);
  return true;
}


bool launch_pair(const ImmaArgs& a, cudaStream_t st) {
  const uint32_t kp = imma_kp(a.k);
  const uint64_t planes = (uint64_t)a.L * a.ell;
  CUtensorMap tmA, tmB;
  if (!make_map(&tmA, a.Mb, {kp, 8, a.rows, planes}, {kp, 8ull * kp, a.Mb_plane}, {KC, 1, RT, 1})) return false;
  // each CTA of the pair loads four of the eight byte planes t of the DT dealers
  if (!make_map(&tmB, a.Vb, {kp, a.Vb_D, 8, planes}, {a.Vb_dstride ? (cuuint64_t)a.Vb_dstride : 8ull * kp, kp, a.Vb_plane}, {KC, 32, 4, 1})) return false;
  const uint32_t b_bytes = 2 * 4 * 32 * KC;
  const uint32_t nstages = std::min<uint32_t>(MAX_STAGES, (SMEM_LIMIT - 1024 - BAR_BYTES - b_bytes) / A_STAGE);
  const uint32_t smem = b_bytes + nstages * A_STAGE + 1024 + BAR_BYTES;
  if (cudaFuncSetAttribute(imma_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) return false;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
  const uint64_t tiles = (uint64_t)((a.rows + 2 * RT - 1) / (2 * RT)) * ((a.D + 31) / 32) * planes;
  if (tiles >= (1ull << 32)) return false;
  const unsigned pairs = (unsigned)std::min<uint64_t>(tiles, (uint64_t)std::max(sms / 2, 1));
  imma_gemm2_kernel<<<2 * pairs, THREADS, smem, st>>>(tmA, tmB, a, nstages);
  return true;
}

}  // namespace

bool imma_shape_ok(uint32_t rows, uint32_t D, uint32_t k) {
  // 8 * k * 255^2 must fit the signed 32-bit accumulator
  return rows > 0 && D > 0 && k >= 1 && k <= 4096;
}

bool launch_imma_gemm(const ImmaArgs& a, cudaStream_t st) {
  if (!imma_shape_ok(a.rows, a.D, a.k)) return false;
  if (a.pair) return launch_pair(a, st);
  return a.dt == 16 ? launch_dt<16>(a, st) : launch_dt<32>(a, st);
}

static bool launch_planes(const u64* M, size_t M_ls, size_t M_rs, uint32_t rows, uint32_t k, uint32_t L, uint32_t ell, uint8_t* Mb, size_t Mb_plane,
                          size_t dst_rs, size_t dst_bs, bool packed, const uint32_t* rowmap, cudaStream_t st) {
  if (rows == 0) return true;
  const uint32_t jc = std::max(4u, 2048u / ell), nchunks = (k + jc - 1) / jc;
  if ((uint64_t)L * nchunks > 65535u) return false;                      // grid.y; L * ceil(k * ell / 2048) is a few hundred at most
  const size_t smem = (size_t)jc * (ell + 1) * 8;
  imma_planes_kernel<<<dim3(rows, L * nchunks), 256, smem, st>>>(M, M_ls, M_rs, k, ell, jc, Mb, Mb_plane, dst_rs, dst_bs, packed ? 1 : 0, rowmap, nchunks);
  return true;
}

bool launch_imma_planes_m(const u64* M, size_t M_ls, size_t M_rs, uint32_t rows, uint32_t k, uint32_t L, uint32_t ell, uint8_t* Mb,
                          size_t Mb_plane, bool packed, cudaStream_t st) {
  const uint32_t kp = imma_kp(k);
  return launch_planes(M, M_ls, M_rs, rows, k, L, ell, Mb, Mb_plane, (size_t)8 * kp, kp, packed, nullptr, st);
}

bool launch_imma_unplanes_m(u64* M, size_t M_ls, size_t M_rs, uint32_t rows, uint32_t k, uint32_t L, uint32_t ell, const uint8_t* Mb, size_t Mb_plane,
                            bool packed, cudaStream_t st) {
  if (rows == 0) return true;
  const uint32_t kp = imma_kp(k), jc = std::max(4u, 2048u / ell), nchunks = (k + jc - 1) / jc;
  if ((uint64_t)L * nchunks > 65535u) return false;
  imma_unplanes_kernel<<<dim3(rows, L * nchunks), 256, (size_t)jc * (ell + 1) * 8, st>>>(M, M_ls, M_rs, k, ell, jc, Mb, Mb_plane, (size_t)8 * kp, kp, packed ? 1 : 0, nchunks);
  return true;
}

bool launch_imma_planes_v(const u64* V, size_t V_ds, size_t V_ls, uint32_t ell, uint32_t D, uint32_t k, uint32_t L, uint8_t* Vb, size_t Vb_plane,
                          bool packed, const uint32_t* dmap, cudaStream_t st) {
  const uint32_t kp = imma_kp(k);
  return launch_planes(V, V_ls, V_ds, D, k, L, ell, Vb, Vb_plane, (size_t)8 * kp, kp, packed, dmap, st);
}

}  // namespace pvw
