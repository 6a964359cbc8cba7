// imma.cuh -- the INT8 tensor-core (tcgen05.mma kind::i8) form of the modular matrix product, see imma.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "modarith.cuh"

namespace pvw {

constexpr uint32_t IMMA_DIAGS = 15;   // byte diagonals of a 8-byte x 8-byte product

// "slot-major" operand form of the tensor-core path: plane = limb*ell + c.
//   M  u64 canonical   M[plane*M_plane + row*k + j]                      (a row = 8k contiguous bytes = the GEMM's K axis)
//   Vx u8              Vx[plane*Vx_plane + (d*15 + u)*8k + 8j + s] = byte (u - s) of V[d][limb][j][c], 0 when out of 0..7
//   O  u64             O[d*O_ds + limb*O_ls + row*O_rs + c*O_cs]        (canonical, or packed halves when O_packed)
//   S  u64 canonical   S[sd*S_ds + limb*S_ls + srow*ell + c],  sd = V_dmap ? V_dmap[d] : d,  srow = S_rowmap ? S_rowmap[row] : row
// mode 0: O = acc + O, 1: O = acc - S, 2: O = acc          (as GemmArgs::mode)
struct ImmaArgs {
  const u64* M; size_t M_plane;
  const uint8_t* Vx; size_t Vx_plane;
  u64* O; size_t O_ls, O_ds, O_rs, O_cs;
  const u64* S; size_t S_ls, S_ds;
  const uint32_t* S_rowmap;
  const uint32_t* V_dmap;
  uint32_t rows, D, k, L, ell;
  int mode, O_packed;
  const LimbConst* lc;   // [L]
};
// false when the shape cannot be served (k odd, tensor-map creation failed): the caller falls back to the IMAD kernel
bool launch_imma_gemm(const ImmaArgs& a, cudaStream_t st);
bool imma_shape_ok(uint32_t rows, uint32_t D, uint32_t k);
// V[sd*V_ds + limb*V_ls + j*ell + c] (canonical, or packed halves)  ->  Vx (layout above), sd = dmap ? dmap[d] : d
void launch_imma_expand(const u64* V, size_t V_ds, size_t V_ls, uint32_t ell, uint32_t D, uint32_t k, uint32_t L, uint8_t* Vx,
                        size_t Vx_plane, bool packed, const uint32_t* dmap, cudaStream_t st);
// limb-major operand (M[limb*M_ls + row*M_rs + j*ell + c], canonical or packed halves) -> slot-major canonical (layout above)
void launch_imma_slot_major(const u64* M, size_t M_ls, size_t M_rs, uint32_t rows, uint32_t k, uint32_t L, uint32_t ell, u64* out,
                            size_t out_plane, bool packed, cudaStream_t st);

}  // namespace pvw
