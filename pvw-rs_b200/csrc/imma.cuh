// imma.cuh -- the INT8 tensor-core (tcgen05.mma kind::i8) form of the modular matrix product, see imma.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "modarith.cuh"

namespace pvw {

constexpr uint32_t IMMA_DIAGS = 15;   // byte diagonals of a 8-byte x 8-byte product
// bytes per byte plane of one operand row: k rounded up to the 16 bytes TMA strides need (the padding is zero)
__host__ __device__ inline uint32_t imma_kp(uint32_t k) { return (k + 15u) & ~15u; }

// Operands of the tensor-core path are BYTE PLANES (plane index p = limb*ell + c):
//   Mb[p*Mb_plane + (row*8 + s)*kp + j] = byte s of M[row][limb][j][c]         (rows of the matrix: A, A^T, B, s_hat)
//   Vb[p*Vb_plane + (d*8 + t)*kp + j] = byte t of V[d][limb][j][c]              (dealer side: r_hat, c1; same form)
//   O  u64   O[d*O_ds + limb*O_ls + row*O_rs + c*O_cs]  canonical, or packed halves when O_packed
//   S  u64   S[sd*S_ds + limb*S_ls + srow*ell + c] canonical, sd = V_dmap ? V_dmap[d] : d, srow = S_rowmap ? S_rowmap[row] : row
// mode 0: O = acc + O, 1: O = acc - S, 2: O = acc          (as GemmArgs::mode)
// Dealers [d_first, d_first + D) of the Vb buffer are multiplied; output / S dealer index d counts from 0.
struct ImmaArgs {
  const uint8_t* Mb; size_t Mb_plane;
  const uint8_t* Vb; size_t Vb_plane; uint32_t Vb_D, d_first;
  size_t Vb_dstride;     // bytes between consecutive dealers of Vb: 0 = 8 * kp (a dense buffer); the ciphertext store's slot stride when
                         // the planes are read in place (then Vb_plane = 8 * kp: a dealer's planes are contiguous)
  u64* O; size_t O_ls, O_ds, O_rs, O_cs;
  const u64* S; size_t S_ls, S_ds;
  const uint32_t* S_rowmap;
  const uint32_t* V_dmap;
  uint32_t rows, D, k, L, ell;
  int mode, O_packed;
  const LimbConst* lc;   // [L]
  int pair;              // 1: the two-SM form (tcgen05.mma.cta_group::2 on CTA pairs) -- measured, not faster; default 0
  int stages;            // depth of the shared-memory ring of M stages: 0 = as deep as fits (the kernel then owns the SM's shared memory);
                         // 2..10 = at most that many, which leaves shared memory for kernels of other streams to co-reside
  int fast_reduce;       // 1: EVERY modulus in lc[0..L) is >= 2^61 (the caller checks): the one-step reduction of the 160-bit sums
  int epi_warps;         // epilogue warps per CTA: 0 / 8 = two per TMEM lane group (default), 16 = four (measured, not faster: imma.cu)
  int dt;                // dealers per tile: 0 = 32 (default), 16 = the half-width tile (probe: tools/csrc/imma_probe.cu)
};
// false when the shape cannot be served (tensor-map creation failed): the caller must have checked imma_shape_ok
bool launch_imma_gemm(const ImmaArgs& a, cudaStream_t st);
bool imma_shape_ok(uint32_t rows, uint32_t D, uint32_t k);
// (both return false when the launch grid cannot be formed)
// limb-major operand M[limb*M_ls + row*M_rs + j*ell + c] (canonical, or packed halves) -> Mb
bool launch_imma_planes_m(const u64* M, size_t M_ls, size_t M_rs, uint32_t rows, uint32_t k, uint32_t L, uint32_t ell, uint8_t* Mb,
                          size_t Mb_plane, bool packed, cudaStream_t st);
// the inverse of launch_imma_planes_m: rebuilds the limb-major operand from its byte planes
bool launch_imma_unplanes_m(u64* M, size_t M_ls, size_t M_rs, uint32_t rows, uint32_t k, uint32_t L, uint32_t ell, const uint8_t* Mb, size_t Mb_plane,
                            bool packed, cudaStream_t st);
// V[sd*V_ds + limb*V_ls + j*ell + c] (canonical, or packed halves), sd = dmap ? dmap[d] : d, d < D  ->  Vb (Vb_D = D)
bool launch_imma_planes_v(const u64* V, size_t V_ds, size_t V_ls, uint32_t ell, uint32_t D, uint32_t k, uint32_t L, uint8_t* Vb,
                          size_t Vb_plane, bool packed, const uint32_t* dmap, cudaStream_t st);

}  // namespace pvw
