// kernels.cuh -- launcher declarations shared by the .cu files of libpvw_b200.so.
// Operand form: every array the multiply-accumulate kernel reads as M or V (A, At, B, r_hat, s_hat, the c1 store) holds each
// residue x < 2^62 as packed 31-bit halves ((x >> 31) << 32 | (x & 0x7fffffff)), so that the inner loop needs no shifts or
// masks; everything else (c2 store, z, decode scratch, all host-facing data) holds canonical residues.
// Device layout ("limb-major"): a vector of polynomials is stored as [..][L][len][ell] so that, for one RNS limb,
// the `len` polynomials' ell-slot blocks are contiguous (len*ell*8 bytes): the modulus is uniform per CTA and a
// row of k polynomials is one contiguous 1D bulk-copy (TMA) source.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "modarith.cuh"

namespace pvw {

// device-resident per-context constant tables
struct DevTables {
  const LimbConst* lc;     // [L]
  const u64* tw;           // [L][ell] psi^brv(i)                     (forward NTT, fhe-math NttOperator omegas)
  const u64* tw_sh;        //          Shoup companions
  const u64* twi;          // [L][ell] psi^-brv(i)                    (inverse NTT)
  const u64* twi_sh;
  const u64* gadget_hat;   // [L][ell] NTT([1, D, .., D^(l-1)] mod q) (parameters.rs:288-308)
  const u64* gadget_hat_sh;  //        Shoup companions
  const u64* dec_c;        // [L][4] multipliers of decode_rns (hostparams.hpp dec_c)
  const u64* lgad;         // [L][ell] ell * Delta^i mod q (power basis, times the unscaled inverse NTT's factor): fused decode check
  const u64* lgad_sh;
  // CRT lift
  const u64* qhat;         // [L][NWT]   Q/q_j, zero padded to the template width
  const u64* Qsh;          // [LB][NWT+1] Q << b
  // decode tail (each NW words): Q, floor(Q/2), M = D^(l-1), floor(M/2), D; then the normalised divisors
  const u64* Qw; const u64* halfQ; const u64* Mw; const u64* halfM; const u64* Dw;
  const u64* divM_v; const u64* div2D_v;
  uint32_t L, ell, NW, NWT, LB;
  uint32_t divM_n, divM_shift, div2D_n, div2D_shift;
  u64 divM_vinv, div2D_vinv;
  // short lift (decode fast path), see hostparams.hpp
  const u64 *sh_c, *sh_c_sh, *sh_qhat, *sh_Q, *sh_halfQ, *sh_v, *sh_v_sh, *sh_r, *sh_r_sh;
  uint32_t shortL, shortSW;
  int lift_fast;  // 1 = try the short lift first
  int tail_impl;  // 1 = register-resident decode tail where a specialisation exists, 0 = generic kernel
  const u64* tern;  // ring degrees 8, 16: [L][l/4][81][l] transforms of the ternary four-coefficient groups (hostparams.hpp), else nullptr
};

// ---- ntt.cu -------------------------------------------------------------------------------------------------
// small signed coefficients -> RNS -> forward NTT (+ optional message encoding), written in the device layout:
//   item idx in [0,count): vec = idx / inner, j = idx % inner;  out[vec*vstride + limb*lstride + j*ell + c]
//   value = NTT(rns(coef[idx]))[c] (+ (m[idx] as i64 mod q) * gadget_hat[limb][c] when m != nullptr)
// pack_out: write the 31-bit-halves operand form (modarith.cuh pack_halves) that the multiply-accumulate kernel reads
// coef: signed integers of `cbytes` bytes each (8 = the reference's i64; 1 / 2 / 4 = the narrow input forms of pvw_b200.h)
// returns false when the shape cannot be launched (ring degree above 256, more than 2^31 blocks) -- nothing was queued then
bool launch_ntt_small(const DevTables& T, const void* coef, int cbytes, const u64* m, uint64_t count, uint32_t inner, u64* out,
                      size_t vstride, size_t lstride, cudaStream_t st, bool accumulate = false, bool pack_out = false, int planes = 0,
                      const u64* addend = nullptr, const void* wide = nullptr, const uint32_t* wide_flag = nullptr);
// wide / wide_flag (byte-plane kernels, k % 4 == 0, ring degree <= 16): `coef` is the one-byte copy launch_narrow_i64 made of the
// 64-bit inputs `wide`; *wide_flag != 0 (a value did not fit one byte) makes the kernel read `wide` instead
bool launch_narrow_i64(const void* in, void* out, uint64_t values, uint32_t* flag, cudaStream_t st);
// true when launch_ntt_small would take the four-polynomial byte-plane kernel for this shape (the one that accepts wide / wide_flag)
inline bool ntt_planes_take_narrow(const DevTables& T, uint32_t inner) { return inner % 4 == 0 && T.ell <= 16; }
// c1 finisher of the tensor-core path (ntt.cu): slot d of the store = [packed residues u64[L][k][ell]][byte planes u8[L*ell][8][kp]];
// addend = the slot-major product [d][limb][c][k].  false: shape not served (k % 4 != 0, ring degree above 16) -- nothing queued.
bool launch_ntt_c1_finish(const DevTables& T, const void* coef, int cbytes, uint64_t count, uint32_t k, u64* c1, size_t slot_stride, const u64* addend,
                          uint32_t kp, cudaStream_t st);
// addend: out = value + addend[(vec*L + limb)*ell*inner + c*inner + j], the slot-major product of the tensor-core kernel
// planes 1 / 2: write the byte planes of the tensor-core product (imma.cuh), matrix-row side (Mb) / dealer side (Vb); then
// inner = k, vstride = kp (bytes per plane row), lstride = plane stride in bytes, out is a byte buffer
// generic strided block copy:  out[b*obs + x*oxs + y*oys + c] = in[b*ibs + x*ixs + y*iys + c],  c < blk
// xform: 0 = plain copy, 1 = canonical residue -> packed halves, 2 = packed halves -> canonical residue
void launch_permute(const u64* in, u64* out, uint64_t Bn, uint64_t X, uint64_t Y, uint32_t blk, size_t ibs, size_t ixs, size_t iys,
                    size_t obs, size_t oxs, size_t oys, cudaStream_t st, int xform = 0);

// ---- mac.cu -------------------------------------------------------------------------------------------------
// NTT-domain polynomial matrix product for every limb and slot:
//   acc[row][d][c] = sum_{j<k} M[limb][row][j][c] * V[d][limb][j][c]  mod q_limb
//   mode 0: O = acc + O (in place: O was pre-loaded with NTT(e) (+ m*g))   -- c1, c2, keygen
//   mode 1: O = acc - S                                                    -- decrypt (S = c2)
//   mode 2: O = acc (store only; NTT(e) (+ m*g) is added afterwards by ntt_small in accumulate mode)
struct GemmArgs {
  const u64* M; size_t M_ls, M_rs;        // M[limb*M_ls + row*M_rs + j*ell + c]
  const u64* V; size_t V_ls, V_ds;        // V[d*V_ds + limb*V_ls + j*ell + c]
  u64* O; size_t O_ls, O_ds;              // O[d*O_ds + limb*O_ls + row*ell + c]
  const u64* S; size_t S_ls, S_ds;        // S[d*S_ds + limb*S_ls + srow(row)*ell + c]
  const uint32_t* S_rowmap;               // optional: srow(row) = S_rowmap[row] (party index list), else row
  const uint32_t* V_dmap;                 // optional: dealer slot list, V/S dealer index = V_dmap[d], else d
  uint32_t rows, D, k, L, ell;
  int mode;
  const LimbConst* lc;
  int O_packed;    // 1: O is written in the packed-halves operand form (it is the M or V operand of a later product)
  int tile;        // register-tile / occupancy variant (mac.cu launch_mac_gemm)
  int refill_lag;  // chunks between a stage's last use and its refill (1 .. NS-1)
};
// impl: 0 = synchronous shared-memory tiles, 1 = cp.async.bulk (TMA) + mbarrier pipeline with a producer warp
// false: the shape cannot be launched (generic kernel: more than 65535 dealers per call)
bool launch_mac_gemm(const GemmArgs& a, int impl, cudaStream_t st);
size_t mac_gemm_launches(const GemmArgs& a);

// ---- decode.cu ----------------------------------------------------------------------------------------------
// share s' = d*Pc + p  (d < D, p < Pc).  z[d*z_ds + limb*z_ls + p*ell + c]  ->  y[(limb*(ell+1) + i)*S + s']
// z_cs == 0: z[d*z_ds + limb*z_ls + p*ell + c]; else the slot-major form z[d*z_ds + limb*z_ls + c*z_cs + p] (tensor-core product).
// sub (optional): the polynomial to subtract first -- z holds <s, c1> only and sub describes c2 (decryption.rs:270-274):
//   S[sd*S_ds + limb*S_ls + srow*ell + c], sd = dmap ? dmap[d] : d, srow = rowmap ? rowmap[p] : p
struct DecodeSub { const u64* S; size_t S_ls, S_ds; const uint32_t* rowmap; const uint32_t* dmap; };
// FallbackList (optional, every kernel of the chain): run only on the shares listed by the fused fast path (device-side count),
// with a grid that does not depend on the count; scratch (y, X) stays indexed by the share number
struct FallbackList { const uint32_t* list; const uint32_t* count; };
bool launch_decode_rns(const DevTables& T, const u64* z, size_t z_ls, size_t z_ds, uint32_t Pc, uint32_t D, u64* y, cudaStream_t st,
                       size_t z_cs = 0, const DecodeSub* sub = nullptr, const FallbackList* fb = nullptr);
// X[(i*NW + w)*S + s'] = CRT lift of y[.][i][s']
void launch_crt_lift(const DevTables& T, const u64* y, u64* X, uint64_t S, cudaStream_t st, const FallbackList* fb = nullptr);
// out[p*out_ps + d] for s' = d*Pc + p
void launch_decode_tail(const DevTables& T, const u64* X, uint32_t Pc, uint32_t D, u64* out, size_t out_ps, cudaStream_t st, const FallbackList* fb = nullptr);
// The fused fast path (decode.cu (0)): clean shares are decoded in one kernel, the others are appended to fb_list (count in
// *fb_count, which the caller zeroes) for the list-driven chain above.  false: not launched (disabled for this parameter set, or
// the shape does not fit) -- the caller then runs the chain on every share.
struct FusedConst {
  u64 dv[4];       // Delta, normalised (top bit of word nd-1 set)
  u64 half_d[4];   // floor(Delta / 2)
  u64 vinv;        // reciprocal of dv[nd-1] (Moller-Granlund)
  u64 cmax;        // largest one-word x with x * Delta^(l-1) + floor(Delta/2) + 1 <= floor(Q/2)
  u64 emax;        // noise magnitudes the claim check takes: l * e < min_j q_j and < 2^63, so that l * e needs no reduction
  uint32_t nd, shift;
  int enabled;
};
bool launch_decode_fused(const DevTables& T, const FusedConst& F, const u64* z, size_t z_ls, size_t z_ds, uint32_t Pc, uint32_t D, u64* out, size_t out_ps,
                         uint32_t* fb_list, uint32_t* fb_count, cudaStream_t st, size_t z_cs = 0, const DecodeSub* sub = nullptr, u64* scr = nullptr);
// scr != nullptr (ring degree 8): two launches -- short lift + carry chain, then the claim check at a higher occupancy -- with
// decode_fused_scratch_words(T, S) words of scratch between them
size_t decode_fused_scratch_words(const DevTables& T, uint64_t S);
size_t decode_scratch_words_y(const DevTables& T, uint64_t S);
size_t decode_scratch_words_X(const DevTables& T, uint64_t S);

// ---- wire.cu ------------------------------------------------------------------------------------------------
// Wire format of one polynomial: a bincode `Vec<u8>` element = u64 length + fhe-math `Poly::to_bytes`
// (protobuf Rq: representation, degree, coefficients = per limb the ell residues packed at bit_length(q_j - 1) bits).
// Every record of one parameter set has the same size and the same first `pre_len` bytes.
struct WireTables {
  const uint8_t* pre;         // [pre_len]  u64 LE message length + the Rq fields before the packed residues
  const uint8_t* limb_of;     // [packed_bytes] limb of each packed byte
  const uint32_t* limb_off;   // [L + 1]    byte offset of each limb's residues inside the packed region
  const uint32_t* nbits;      // [L]        bits per residue
  const uint32_t* magic;      // [L]        floor(2^20 / nbits) + 1:  x / nbits == (x * magic) >> 20 for x < 2^11
  const u64* moduli;          // [L]
  uint32_t L, ell, pre_len, packed_bytes, rec_bytes;
  uint32_t sg;                // lanes per (record, limb) pair in the pack kernel: 16 when every limb has <= 16 words, else 32
};
enum { WIRE_ERR_HEADER = 1, WIRE_ERR_RESIDUE = 2, WIRE_ERR_ENVELOPE = 4 };
// polynomials src[b*bs + limb*ls + p*ell + c]  (p < count, b < batch)  ->  records out[b*obs + p*rec_bytes ..]
void launch_wire_pack(const WireTables& W, const u64* src, size_t ls, size_t bs, uint64_t count, uint32_t batch, int src_packed,
                      uint8_t* out, size_t obs, cudaStream_t st);
// the inverse; write == 0 only validates.  *err |= WIRE_ERR_* on a malformed record / a residue >= q_j
void launch_wire_unpack(const WireTables& W, const uint8_t* in, size_t ibs, uint64_t count, uint32_t batch, u64* dst, size_t ls, size_t bs,
                        int dst_packed, int write, int* err, cudaStream_t st);
// out[b*obs + i] = tmpl[i]  /  *err |= WIRE_ERR_ENVELOPE when in[b*ibs + i] != tmpl[i]      (i < len, b < batch)
void launch_wire_fill(uint8_t* out, size_t obs, uint32_t batch, const uint8_t* tmpl, uint32_t len, cudaStream_t st);
void launch_wire_expect(const uint8_t* in, size_t ibs, uint32_t batch, const uint8_t* tmpl, uint32_t len, int* err, cudaStream_t st);

}  // namespace pvw
