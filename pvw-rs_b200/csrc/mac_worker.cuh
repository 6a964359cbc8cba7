// mac_worker.cuh -- the register-tile inner loop of the matrix multiply-accumulate kernel (mac.cu), shared with the
// integer-pipe microbenchmark (tools/csrc/int_peaks.cu) so that tile shapes can be timed without the copy pipeline.
#pragma once
#include "kernels.cuh"


namespace pvw {

constexpr int kComputeThreads = 256;
constexpr int kStages = 4;  // default pipeline depth

template <int ELL, int TR, int TD, int GD, int KC, int THREADS = kComputeThreads, bool DENSE_M = false>
struct TileCfg {
  // KC = polynomials (j indices) per pipeline stage
  static constexpr int G = THREADS / ELL;  // (row-group, dealer-group) pairs per CTA
  static constexpr int GR = G / GD;
  static constexpr int RT = GR * TR;               // rows per CTA
  static constexpr int DT = GD * TD;               // dealers per CTA
  // bytes per staged row: the pad staggers the (32/ELL) sub-tiles of a warp that read DIFFERENT rows (dealer sub-tiles
  // when GD > 1, row sub-tiles otherwise) onto disjoint banks: a half-warp of 8-byte loads covers 128 bytes, so two
  // neighbouring sub-tiles must sit 64 bytes apart modulo 128
  static constexpr int STEP = GD > 1 ? TD : TR;  // staged rows between neighbouring sub-tiles of a warp
  static constexpr int pick_pad() {
    if (ELL != 8) return 16;
    for (int pad = 16; pad <= 128; pad += 16)
      if ((STEP * (KC * ELL * 8 + pad)) % 128 == 64) return pad;
    return 16;
  }
  static constexpr int ROWB = KC * ELL * 8 + pick_pad();
  // DENSE_M: the matrix rows arrive as ONE tensor-map (TMA) box, densely packed (pitch = KC*ELL*8); sub-tiles that share
  // a row read it as a broadcast, so no stagger is needed there -- only the dealer rows keep the padded pitch
  static constexpr int MROWB = DENSE_M ? KC * ELL * 8 : ROWB;
  static constexpr int VROWB = ROWB;
  static constexpr int MBYTES = RT * MROWB;
  static constexpr int STAGE = ((MBYTES + DT * VROWB + 127) / 128) * 128;
  static_assert(G % GD == 0, "bad tile");
};

// NJ_: polynomials per straight-line block; ROLL: blocks in a real loop (small code) instead of unrolled;
// PACKED: operands are stored as 31-bit halves (x1 << 32 | x0, modarith.cuh pack_halves) instead of canonical residues
template <int ELL, int TR, int TD, int GD, int KC, int NJ_ = 4, bool ROLL = false, bool PACKED = true, int THREADS_ = kComputeThreads,
          bool DENSE_M = false>
struct Worker {
  using C = TileCfg<ELL, TR, TD, GD, KC, THREADS_, DENSE_M>;
  static constexpr int THREADS = THREADS_;
  int c, gr, gd;
  u32 zero;  // run-time 0 (see add_alu)
  AccK acc[TR][TD];
  __device__ __forceinline__ void init(int tid, u32 opaque_zero = 0) {
    zero = opaque_zero;
    const int lane = tid & 31, w = tid >> 5;
    c = lane % ELL;
    const int g = w * (32 / ELL) + lane / ELL;
    gd = g % GD;
    gr = g / GD;
#pragma unroll
    for (int t = 0; t < TR; t++)
#pragma unroll
      for (int u = 0; u < TD; u++) acck_zero(acc[t][u]);
  }
  // NJ consecutive polynomials starting at jj0: the TD dealer operands are split once and reused for the TR rows
  template <int NJ>
  __device__ __forceinline__ void block(const unsigned char* ms, const unsigned char* vs, int jj0) {
    u32 b0[TD][NJ], b1[TD][NJ];  // 31-bit halves; the Karatsuba sum is formed at the use (alu pipe has slack, registers do not)
#pragma unroll
    for (int u = 0; u < TD; u++)
#pragma unroll
      for (int i = 0; i < NJ; i++) {
        const u64 x = *reinterpret_cast<const u64*>(vs + u * C::VROWB + (jj0 + i) * ELL * 8);
        b0[u][i] = PACKED ? (u32)x : ((u32)x & 0x7fffffffu);
        b1[u][i] = PACKED ? (u32)(x >> 32) : (u32)(x >> 31);
      }
#pragma unroll
    for (int t = 0; t < TR; t++) {
      SplitOp a[NJ];
#pragma unroll
      for (int i = 0; i < NJ; i++) {
        const u64 x = *reinterpret_cast<const u64*>(ms + t * C::MROWB + (jj0 + i) * ELL * 8);
        if (PACKED) { a[i].x0 = (u32)x; a[i].x1 = (u32)(x >> 32); a[i].xs = add_alu(a[i].x0, a[i].x1, zero); }
        else a[i] = split_op(x, zero);
      }
#pragma unroll
      for (int u = 0; u < TD; u++)
#pragma unroll
        for (int i = 0; i < NJ; i++) {  // (loop order is irrelevant: ptxas re-schedules the independent carry chains)
          SplitOp b;
          b.x0 = b0[u][i]; b.x1 = b1[u][i]; b.xs = add_alu(b0[u][i], b1[u][i], zero);
          acck_mac(acc[t][u], a[i], b);
        }
    }
  }
  // one staged chunk: kc polynomials of every row / dealer of the tile
  template <bool FULL>
  __device__ __forceinline__ void chunk(const unsigned char* stage, int kc) {
    const unsigned char* ms = stage + (size_t)(gr * TR) * C::MROWB + c * 8;
    const unsigned char* vs = stage + C::MBYTES + (size_t)(gd * TD) * C::VROWB + c * 8;
    constexpr int NJ = KC % NJ_ == 0 ? NJ_ : (KC % 2 == 0 ? 2 : 1);
    if (FULL && ROLL) {
#pragma unroll 1
      for (int jj = 0; jj < KC; jj += NJ) block<NJ>(ms, vs, jj);
    } else if (FULL) {
      // straight-line blocks of NJ polynomials; the compiler barrier keeps ptxas from hoisting the next block's
      // shared-memory loads above this block's arithmetic (which costs registers and then spills)
#pragma unroll
      for (int jj = 0; jj < KC; jj += NJ) {
        block<NJ>(ms, vs, jj);
#ifndef PVW_EXP_NO_BARRIER  // measured: without the barrier the kernel spills and runs at half the rate
        asm volatile("" ::: "memory");
#endif
      }
    } else {
#pragma unroll 1
      for (int jj = 0; jj < kc; jj++) block<1>(ms, vs, jj);
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
        u64 v = acck_reduce(acc[t][u], lc);
        u64* o = g.O + (size_t)d * g.O_ds + (size_t)limb * g.O_ls + (size_t)row * ELL + c;
        if (g.mode == 0) {
          v = addmod(v, *o, lc.q);
        } else if (g.mode == 2) {
        } else {
          const uint32_t srow = g.S_rowmap ? g.S_rowmap[row] : row;
          v = submod(v, g.S[(size_t)ds * g.S_ds + (size_t)limb * g.S_ls + (size_t)srow * ELL + c], lc.q);
        }
        *o = g.O_packed ? pack_halves(v) : v;
      }
    }
  }
};

}  // namespace pvw
