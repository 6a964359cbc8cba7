// capi.cu -- the C ABI of libpvw_b200.so (include/pvw_b200.h): context, device residency of A / B / ciphertexts
// and the host-side sequencing of the kernels in ntt.cu, mac.cu and decode.cu.
//
// Device layout ("limb-major", see kernels.cuh):
//   A    u64[L][k][k][ell]          A[i][j] at ((limb*k + i)*k + j)*ell          PvwCrs.matrix          crs.rs:12-17
//   At   u64[L][k][k][ell]          At[c][j] = A[j][c] (built lazily for keygen)  multiply_by_secret_key crs.rs:152-165
//   B    u64[L][nrows][k][ell]      local rows of GlobalPublicKey.matrix           public_key.rs:43-54
//   c1   u64[cap][L][k][ell]        ciphertext store, PvwCiphertext.c1             encryption.rs:15-24
//   c2   u64[cap][L][nrows][ell]                      PvwCiphertext.c2 (local rows)
// Host layout at the boundary is always the reference's: polynomial = u64[L][ell] row-major.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/pvw_b200.h"
#include "hostparams.hpp"
#include "crsgen.hpp"
#include "kernels.cuh"
#include "imma.cuh"

using namespace pvw;

namespace {

thread_local std::string g_create_error;

std::string fmt(const char* f, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return std::string(buf);
}

#define CUDA_CHECK(expr)                                                                                        \
  do {                                                                                                          \
    cudaError_t e__ = (expr);                                                                                   \
    if (e__ != cudaSuccess)                                                                                     \
      throw PvwException(PVW_ERR_INTERNAL, fmt("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #expr)); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  void ensure(size_t need) {
    if (need <= bytes) return;
    if (p) CUDA_CHECK(cudaFree(p));
    p = nullptr; bytes = 0;
    CUDA_CHECK(cudaMalloc(&p, need));
    bytes = need;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct pvw_ctx {
  HostParams hp;
  int device = 0;
  uint32_t row0 = 0, nrows = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // host<->device staging copies of the host-pointer calls, overlapped with kernels
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // [0] e2 / m staged, [1] r / e1 staged, [2] tail of the last host-pointer encryption, [3] a decrypt chunk done
  bool tail_pending = false;           // ev[2] has been recorded: the copy stream of a later host-pointer call must wait for it
  std::vector<cudaEvent_t> chunk_ev;
  DevTables T{};
  FusedConst F{};                // decode fast path constants (decode.cu (0))
  int decode_fused = 1;          // option "decode_fused": 0 = always the three-kernel chain, 1 = the fused fast path first
  DevBuf fb;                     // [count (16 bytes)][list u32[S]] of the shares the fast path hands to the chain
  DevBuf tables;                 // one allocation holding every constant table
  DevBuf A, At, B;
  bool A_set = false, At_valid = false;
  uint32_t num_keys = 0;         // max uploaded global index + 1 (public_key.rs:245-247)
  uint32_t cap = 0;
  DevBuf c1s, c2s;               // c1s: per slot [packed residues u64[L][k][ell]][byte planes u8[L*ell][8][kp]] (c1_stride() words)
  std::vector<uint8_t> c1p_valid;  // per slot: the byte planes behind the residues are current (written by the c1 finisher / a peer's push)
  bool c1_external = false;        // the raw store pointer was handed out (pvw_ct_c1_device_ptr): planes can go stale behind our back
  DevBuf prod1;                    // slot-major c1 product of one encrypt call
  DevBuf narrow;                   // one-byte copy of a chunk of 64-bit secrets + 64 flag words (ntt.cu narrow_i64_kernel)
  DevBuf fscr;                     // fused decode, two-launch form: l + 2 words per share between the launches
  // grow-only scratch
  DevBuf stage, rhat, in_small, in_small2, in_m, shat, z, y, X, outd, idxd, idxp;
  // wire format (wire.cu): record tables, envelope templates (device copies at wire_env) and a staging buffer
  WireTables W{};
  DevBuf wire_tab, wire_buf;
  std::vector<uint8_t> params_blob;                  // bincode(PvwParameters), parameters.rs:606-623
  const uint8_t *env_k = nullptr, *env_n = nullptr, *env_params = nullptr;
  int* wire_err = nullptr;
  uint64_t rq_bytes = 0;
  std::vector<uint32_t> h_idxp;  // local party indices of the current decrypt call
  std::string err;
  uint64_t launches = 0;
  // optional per-kernel-kind CUDA-event timing (bench.py's roofline leg): events bracket each launch on `stream`
  bool profile = false;
  struct ProfRec { int kind; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[PVW_KERNEL_KINDS] = {0};
  uint64_t prof_n[PVW_KERNEL_KINDS] = {0};
  double prof_bytes[PVW_KERNEL_KINDS] = {0};
  int gemm_impl = 1, gemm_tile = 1, refill_lag = 2;
  // tensor-core form of the matrix product (imma.cu): slot-major canonical copies of A / B (built lazily from the operand
  // form), the diagonal expansion of the dealer-side operand, the slot-major secret-key transforms of one decrypt chunk
  int use_tern = 1, narrow_inputs = 1;
  const u64* tern_dev = nullptr;
  int use_imma = 1, imma_pair = 0, imma_stages = 0, imma_epi_warps = 0, imma_fast_reduce = 1;
  int64_t imma_min_dealers = 8, imma_min_rows = 16, imma_chunk_dealers = 512;
  DevBuf As, Bs, Vx, shat_s, prod;
  bool As_valid = false, Bs_valid = false;
  int planes_only = 0;           // option "planes_only": once the byte planes of B exist, free its u64 copy (rebuilt on demand)
  bool B_released = false;
  int64_t decrypt_chunk_shares = 1 << 19;
  int64_t upload_chunk_bytes = 256ll << 20;
  // multi-GPU c1 exchange over the copy engines (pvw_shard_*): peers' ciphertext stores and flag arrays mapped through CUDA IPC
  struct Shard {
    uint32_t world = 0, rank = 0;
    std::vector<u64*> peer_c1;          // [world] IPC mappings of the peers' c1 stores (own entry: the local store)
    std::vector<u64*> peer_flags;       // [world] IPC mappings of the peers' flag arrays
    u64* flags = nullptr;               // local: [0, world) data counters written by the peers' pushes, [world, 2 world) their acks,
                                        //        [2 world], [2 world + 1] staging words for the values this rank sends
    uint32_t flags_world = 0;           // the world size `flags` was allocated for
    u64 push_seq = 0, wait_seq = 0;
    uint32_t batch_slot0 = 0, batch_D = 0;   // the encrypt call whose c1 slice was pushed last (PVW_ENC_PUSH_C1) ...
    bool batch_direct = false;               // ... and whether its slots carry current byte planes (every rank takes the same path)
    cudaStream_t xstream = nullptr;     // the exchange stream (copy engines only)
    cudaEvent_t ev_c1 = nullptr;
    bool connected = false;
  } sh;

  size_t poly() const { return (size_t)hp.L * hp.ell; }
  size_t c1_words() const { return (size_t)hp.L * hp.k * hp.ell; }                                   // residues of one c1
  size_t c1_stride() const { return c1_words() + (size_t)hp.L * hp.ell * imma_kp(hp.k); }            // + its byte planes, in u64 words
  void c1_planes_invalidate(uint32_t slot0, uint32_t count) { for (uint32_t i = slot0; i < slot0 + count && i < c1p_valid.size(); i++) c1p_valid[i] = 0; }
  void use() { CUDA_CHECK(cudaSetDevice(device)); }
};

namespace {

// ------------------------------------------------------------------------------------------------------------
// constant tables -> device
// ------------------------------------------------------------------------------------------------------------
uint32_t lift_template_width(uint32_t nw) {
  static const uint32_t widths[] = {2, 4, 8, 17, 33, 64};
  for (uint32_t w : widths) if (nw <= w) return w;
  throw PvwException(PVW_ERR_INVALID_PARAMETERS, "modulus product wider than 4096 bits is not supported");
}

void upload_tables(pvw_ctx* c) {
  const HostParams& hp = c->hp;
  const uint32_t L = hp.L, ell = hp.ell, NW = hp.NW, NWT = lift_template_width(NW), LB = hp.LB;
  std::vector<uint64_t> blob;
  auto put = [&](const uint64_t* src, size_t n) { size_t off = blob.size(); blob.insert(blob.end(), src, src + n); return off; };
  auto pad2 = [&]() { if (blob.size() & 1) blob.push_back(0); };
  static_assert(sizeof(LimbConst) % 8 == 0, "LimbConst must be a whole number of words");
  size_t o_lc = put(reinterpret_cast<const uint64_t*>(hp.lc.data()), (size_t)L * sizeof(LimbConst) / 8); pad2();
  size_t o_tw = put(hp.tw.data(), (size_t)L * ell), o_tws = put(hp.tw_sh.data(), (size_t)L * ell);
  size_t o_twi = put(hp.twi.data(), (size_t)L * ell), o_twis = put(hp.twi_sh.data(), (size_t)L * ell);
  size_t o_g = put(hp.gadget_hat.data(), (size_t)L * ell), o_gs = put(hp.gadget_hat_sh.data(), (size_t)L * ell);
  size_t o_dc = put(hp.dec_c.data(), (size_t)L * 4);
  size_t o_lg = put(hp.lgad.data(), (size_t)L * ell), o_lgs = put(hp.lgad_sh.data(), (size_t)L * ell);
  std::vector<uint64_t> qhat_t((size_t)L * NWT, 0), qsh_t((size_t)LB * (NWT + 1), 0);
  for (uint32_t j = 0; j < L; j++) memcpy(&qhat_t[(size_t)j * NWT], &hp.qhat[(size_t)j * NW], (size_t)NW * 8);
  for (uint32_t b = 0; b < LB; b++) memcpy(&qsh_t[(size_t)b * (NWT + 1)], &hp.Qsh[(size_t)b * (NW + 1)], (size_t)(NW + 1) * 8);
  size_t o_qhat = put(qhat_t.data(), qhat_t.size()), o_qsh = put(qsh_t.data(), qsh_t.size());
  size_t o_Q = put(hp.Qw.data(), NW), o_hQ = put(hp.halfQ.data(), NW), o_M = put(hp.Mw.data(), NW), o_hM = put(hp.halfM.data(), NW),
         o_D = put(hp.Dw.data(), NW);
  std::vector<uint64_t> dM(hp.divM.v), d2D(hp.div2D.v);
  size_t o_dM = put(dM.data(), dM.size()), o_d2D = put(d2D.data(), d2D.size());
  auto putv = [&](const std::vector<uint64_t>& v) { pad2(); return v.empty() ? blob.size() : put(v.data(), v.size()); };
  size_t o_shc = putv(hp.sh_c), o_shcs = putv(hp.sh_c_sh), o_shq = putv(hp.sh_qhat), o_shQ = putv(hp.sh_Q), o_shh = putv(hp.sh_halfQ),
         o_shv = putv(hp.sh_v), o_shvs = putv(hp.sh_v_sh), o_shr = putv(hp.sh_r), o_shrs = putv(hp.sh_r_sh), o_tern = putv(hp.tern);
  c->tables.ensure(blob.size() * 8);
  CUDA_CHECK(cudaMemcpy(c->tables.p, blob.data(), blob.size() * 8, cudaMemcpyHostToDevice));
  const u64* base = c->tables.as<u64>();
  DevTables& T = c->T;
  T.lc = reinterpret_cast<const LimbConst*>(base + o_lc);
  T.tw = base + o_tw; T.tw_sh = base + o_tws; T.twi = base + o_twi; T.twi_sh = base + o_twis; T.gadget_hat = base + o_g; T.gadget_hat_sh = base + o_gs; T.dec_c = base + o_dc;
  T.lgad = base + o_lg; T.lgad_sh = base + o_lgs;
  T.qhat = base + o_qhat; T.Qsh = base + o_qsh;
  T.Qw = base + o_Q; T.halfQ = base + o_hQ; T.Mw = base + o_M; T.halfM = base + o_hM; T.Dw = base + o_D;
  T.divM_v = base + o_dM; T.div2D_v = base + o_d2D;
  T.L = L; T.ell = ell; T.NW = NW; T.NWT = NWT; T.LB = LB;
  T.divM_n = hp.divM.n; T.divM_shift = hp.divM.shift; T.divM_vinv = hp.divM.vinv;
  T.div2D_n = hp.div2D.n; T.div2D_shift = hp.div2D.shift; T.div2D_vinv = hp.div2D.vinv;
  T.tail_impl = 1;
  T.sh_c = base + o_shc; T.sh_c_sh = base + o_shcs; T.sh_qhat = base + o_shq; T.sh_Q = base + o_shQ; T.sh_halfQ = base + o_shh;
  T.sh_v = base + o_shv; T.sh_v_sh = base + o_shvs; T.sh_r = base + o_shr; T.sh_r_sh = base + o_shrs;
  c->tern_dev = hp.tern.empty() ? nullptr : base + o_tern;
  T.tern = c->use_tern ? c->tern_dev : nullptr;
  T.shortL = hp.shortL; T.shortSW = hp.shortSW; T.lift_fast = hp.shortL > 0 ? 1 : 0;
  FusedConst& F = c->F;
  memset(&F, 0, sizeof(F));
  F.enabled = hp.fused_ok ? 1 : 0;
  if (hp.fused_ok) {
    F.nd = hp.divD.n; F.shift = hp.divD.shift; F.vinv = hp.divD.vinv; F.cmax = hp.fused_cmax; F.emax = hp.fused_emax;
    for (uint32_t i = 0; i < 4; i++) { F.dv[i] = i < hp.divD.v.size() ? hp.divD.v[i] : 0; F.half_d[i] = hp.half_delta[i]; }
  }
}

// ------------------------------------------------------------------------------------------------------------
// wire format constants (recalled third-party encodings, see oracle/pvw_wire.py)
// ------------------------------------------------------------------------------------------------------------
void put_varint(std::vector<uint8_t>& v, uint64_t x) {
  while (x >= 0x80) { v.push_back((uint8_t)(x | 0x80)); x >>= 7; }
  v.push_back((uint8_t)x);
}
void put_u64le(std::vector<uint8_t>& v, uint64_t x) { for (int i = 0; i < 8; i++) v.push_back((uint8_t)(x >> (8 * i))); }

void upload_wire_tables(pvw_ctx* c) {
  const HostParams& hp = c->hp;
  const uint32_t L = hp.L, ell = hp.ell;
  std::vector<uint32_t> nbits(L), magic(L), limb_off(L + 1, 0);
  for (uint32_t j = 0; j < L; j++) {
    uint32_t nb = 0;
    for (uint64_t x = hp.moduli[j] - 1; x; x >>= 1) nb++;                 // fhe-math Modulus::serialize_vec: 64 - lzcnt(p - 1)
    nbits[j] = nb;
    magic[j] = (1u << 20) / nb + 1;
    limb_off[j + 1] = limb_off[j] + ell * nb / 8;                         // ell is a multiple of 8: no padding bits
  }
  const uint32_t packed = limb_off[L];
  std::vector<uint8_t> limb_of(packed);
  for (uint32_t j = 0; j < L; j++) std::fill(limb_of.begin() + limb_off[j], limb_of.begin() + limb_off[j + 1], (uint8_t)j);
  // Rq { representation = NTT (2); degree = ell; coefficients = packed; allow_variable_time = false (omitted) }
  std::vector<uint8_t> rq = {0x08, 0x02, 0x10};
  put_varint(rq, ell);
  rq.push_back(0x1a);
  put_varint(rq, packed);
  c->rq_bytes = rq.size() + packed;
  std::vector<uint8_t> pre;
  put_u64le(pre, c->rq_bytes);                                            // bincode Vec<u8> length
  pre.insert(pre.end(), rq.begin(), rq.end());
  // bincode(PvwParameters): n, k, l (usize as u64), moduli Vec<u64>, secret_variance f32, the bounds as decimal strings
  std::vector<uint8_t>& pb = c->params_blob;
  pb.clear();
  put_u64le(pb, hp.n); put_u64le(pb, hp.k); put_u64le(pb, hp.ell);
  put_u64le(pb, L);
  for (uint32_t j = 0; j < L; j++) put_u64le(pb, hp.moduli[j]);
  uint32_t fbits; memcpy(&fbits, &hp.secret_variance, 4);
  for (int i = 0; i < 4; i++) pb.push_back((uint8_t)(fbits >> (8 * i)));
  for (uint64_t b : {hp.b1, hp.b2}) {
    const std::string sdec = std::to_string(b);
    put_u64le(pb, sdec.size());
    pb.insert(pb.end(), sdec.begin(), sdec.end());
  }
  // one device blob: [u32 tables][u64 moduli][bytes]
  std::vector<uint8_t> blob;
  auto put = [&](const void* src, size_t bytes, size_t align) {
    while (blob.size() % align) blob.push_back(0);
    const size_t off = blob.size();
    blob.insert(blob.end(), (const uint8_t*)src, (const uint8_t*)src + bytes);
    return off;
  };
  const size_t o_mod = put(hp.moduli.data(), (size_t)L * 8, 8), o_nb = put(nbits.data(), (size_t)L * 4, 4), o_mg = put(magic.data(), (size_t)L * 4, 4),
               o_lo = put(limb_off.data(), (size_t)(L + 1) * 4, 4), o_lf = put(limb_of.data(), packed, 4), o_pre = put(pre.data(), pre.size(), 4);
  std::vector<uint8_t> ek, en;
  put_u64le(ek, hp.k); put_u64le(en, hp.n);
  const size_t o_ek = put(ek.data(), 8, 8), o_en = put(en.data(), 8, 8), o_pb = put(pb.data(), pb.size(), 8);
  int zero = 0;
  const size_t o_err = put(&zero, 4, 4);
  c->wire_tab.ensure(blob.size());
  CUDA_CHECK(cudaMemcpy(c->wire_tab.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  uint8_t* base = c->wire_tab.as<uint8_t>();
  WireTables& W = c->W;
  W.moduli = reinterpret_cast<const u64*>(base + o_mod);
  W.nbits = reinterpret_cast<const uint32_t*>(base + o_nb);
  W.magic = reinterpret_cast<const uint32_t*>(base + o_mg);
  W.limb_off = reinterpret_cast<const uint32_t*>(base + o_lo);
  W.limb_of = base + o_lf;
  W.pre = base + o_pre;
  W.L = L; W.ell = ell; W.pre_len = (uint32_t)pre.size(); W.packed_bytes = packed; W.rec_bytes = (uint32_t)pre.size() + packed;
  W.sg = 16;
  for (uint32_t j = 0; j < L; j++) if (limb_off[j + 1] - limb_off[j] > 64) W.sg = 32;
  c->env_k = base + o_ek; c->env_n = base + o_en; c->env_params = base + o_pb;
  c->wire_err = reinterpret_cast<int*>(base + o_err);
}

void check_launch(pvw_ctx* c, size_t n = 1) {
  c->launches += n;
  CUDA_CHECK(cudaGetLastError());
}

cudaEvent_t prof_event(pvw_ctx* c) {
  if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
  cudaEvent_t e;
  CUDA_CHECK(cudaEventCreate(&e));
  return e;
}
void prof_drain(pvw_ctx* c) {
  if (c->prof_pending.empty()) return;
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (auto& r : c->prof_pending) {
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
    c->prof_ms[r.kind] += ms;
    c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b);
  }
  c->prof_pending.clear();
}
// run one kernel launch `f`, counting it and (when profiling) bracketing it with events; `bytes` = algorithmic bytes
template <class F>
void launch(pvw_ctx* c, int kind, double bytes, F&& f) {
  if (!c->profile) { f(); check_launch(c); return; }
  cudaEvent_t a = prof_event(c), b = prof_event(c);
  CUDA_CHECK(cudaEventRecord(a, c->stream));
  f();
  check_launch(c);
  CUDA_CHECK(cudaEventRecord(b, c->stream));
  c->prof_pending.push_back({kind, a, b});
  c->prof_n[kind]++;
  c->prof_bytes[kind] += bytes;
  if (c->prof_pending.size() >= 4096) prof_drain(c);
}

// copy `bytes` from a caller pointer (host or device, per flags) into a device scratch buffer
const void* stage_in(pvw_ctx* c, DevBuf& buf, const void* src, size_t bytes, uint32_t flags) {
  if (flags & PVW_IO_DEVICE) return src;
  buf.ensure(bytes);
  CUDA_CHECK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return buf.p;
}

// host-pointer calls: the same copy, but on the copy stream, so that it overlaps kernels already queued on the compute
// stream; `ready` is recorded after it and the consumer makes the compute stream wait for it
const void* stage_in_overlapped(pvw_ctx* c, DevBuf& buf, const void* src, size_t bytes, uint32_t flags, cudaEvent_t ready) {
  if (flags & PVW_IO_DEVICE) return src;
  buf.ensure(bytes);
  CUDA_CHECK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
  CUDA_CHECK(cudaEventRecord(ready, c->copy_stream));
  return buf.p;
}

// host-layout polynomials [count][L][ell]  <->  limb-major [L][count][ell] slice inside a bigger array
void to_limb_major(pvw_ctx* c, const u64* in_host_layout, uint64_t count, u64* out, size_t out_limb_stride, bool operand) {
  const uint32_t L = c->hp.L, ell = c->hp.ell;
  launch(c, PVW_KERNEL_PERMUTE, 0.0, [&] { launch_permute(in_host_layout, out, 1, count, L, ell, 0, (size_t)L * ell, ell, 0, ell, out_limb_stride, c->stream, operand ? 1 : 0); });
}
void from_limb_major(pvw_ctx* c, const u64* in, size_t in_limb_stride, uint64_t count, u64* out_host_layout, bool operand) {
  const uint32_t L = c->hp.L, ell = c->hp.ell;
  // x = limb is the fastest thread axis here so that the host-layout side (the output) is written contiguously
  launch(c, PVW_KERNEL_PERMUTE, 0.0, [&] { launch_permute(in, out_host_layout, 1, L, count, ell, 0, in_limb_stride, ell, 0, ell, (size_t)L * ell, c->stream, operand ? 2 : 0); });
}

// upload `count` polynomials given in host layout (host or device memory) into limb-major dst (+ offset handled by caller)
void upload_polys(pvw_ctx* c, const uint64_t* src, uint64_t count, u64* dst, size_t dst_limb_stride, uint32_t flags, bool operand) {
  const size_t poly = c->poly();
  if (flags & PVW_IO_DEVICE) { to_limb_major(c, reinterpret_cast<const u64*>(src), count, dst, dst_limb_stride, operand); return; }
  uint64_t per = std::max<uint64_t>(1, (uint64_t)c->upload_chunk_bytes / (poly * 8));
  for (uint64_t o = 0; o < count; o += per) {
    uint64_t n = std::min(per, count - o);
    c->stage.ensure(n * poly * 8);
    CUDA_CHECK(cudaMemcpyAsync(c->stage.p, src + o * poly, n * poly * 8, cudaMemcpyHostToDevice, c->stream));
    to_limb_major(c, c->stage.as<u64>(), n, dst + o * c->hp.ell, dst_limb_stride, operand);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));  // the staging buffer is reused by the next chunk
  }
}
void download_polys(pvw_ctx* c, const u64* src, size_t src_limb_stride, uint64_t count, uint64_t* dst_host, bool operand) {
  const size_t poly = c->poly();
  uint64_t per = std::max<uint64_t>(1, (uint64_t)c->upload_chunk_bytes / (poly * 8));
  for (uint64_t o = 0; o < count; o += per) {
    uint64_t n = std::min(per, count - o);
    c->stage.ensure(n * poly * 8);
    from_limb_major(c, src + o * c->hp.ell, src_limb_stride, n, c->stage.as<u64>(), operand);
    CUDA_CHECK(cudaMemcpyAsync(dst_host + o * poly, c->stage.p, n * poly * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
}

void require(bool cond, int code, const std::string& msg) { if (!cond) throw PvwException(code, msg); }

// element size of the small signed inputs of a call (pvw_b200.h PVW_IN_*): secrets / randomness, and errors
int secret_bytes(uint32_t flags) { return (flags & PVW_IN_SECRET_I8) ? 1 : 8; }
int error_bytes(uint32_t flags) {
  require(!((flags & PVW_IN_ERROR_I32) && (flags & PVW_IN_ERROR_I16)), PVW_ERR_INVALID_PARAMETERS, "PVW_IN_ERROR_I32 and PVW_IN_ERROR_I16 exclude each other");
  return (flags & PVW_IN_ERROR_I32) ? 4 : (flags & PVW_IN_ERROR_I16) ? 2 : 8;
}
const void* at_bytes(const void* p, size_t bytes) { return reinterpret_cast<const uint8_t*>(p) + bytes; }

// ntt.cu: small signed coefficients (element size cbytes) -> NTT form in one of the device layouts
void ntt(pvw_ctx* c, const void* coef, int cbytes, const u64* m, uint64_t count, uint32_t inner, u64* out, size_t vstride, size_t lstride,
         bool accumulate = false, bool pack_out = false, int planes = 0, const u64* addend = nullptr, const void* wide = nullptr,
         const uint32_t* wide_flag = nullptr) {
  bool ok = true;
  launch(c, planes ? PVW_KERNEL_NTT_PLANES : PVW_KERNEL_NTT, 0.0, [&] { ok = launch_ntt_small(c->T, coef, cbytes, m, count, inner, out, vstride, lstride, c->stream, accumulate, pack_out, planes, addend, wide, wide_flag); });
  require(ok, PVW_ERR_INTERNAL, "forward NTT: unsupported shape (ring degree above 256 or more than 2^31 thread blocks)");
}

void gemm(pvw_ctx* c, GemmArgs a) {
  a.tile = c->gemm_tile;
  a.refill_lag = a.D == 1 ? 1 : c->refill_lag;  // the HBM-bound matrix-vector form wants the deepest prefetch
  // algorithmic bytes: per (dealer, row) one k-polynomial operand row read + one polynomial written
  const double bytes = (double)a.D * a.rows * (a.k + 1.0) * a.L * a.ell * 8.0;
  bool ok = true;
  launch(c, PVW_KERNEL_MAC, bytes, [&] { ok = launch_mac_gemm(a, c->gemm_impl, c->stream); });
  require(ok, PVW_ERR_INTERNAL, "matrix product: unsupported shape");
}

// ---- tensor-core product (imma.cu) --------------------------------------------------------------------------------
bool imma_wanted(const pvw_ctx* c, uint32_t rows, uint32_t D) {
  // a tile is 128 rows x 32 dealers: below ~16 rows (e.g. one party decrypting many ciphertexts) or 8 dealers the CUDA-core
  // kernel, which streams the big operand once, is the faster one
  return c->use_imma && D >= (uint32_t)c->imma_min_dealers && rows >= (uint32_t)c->imma_min_rows && imma_shape_ok(rows, D, c->hp.k);
}
// zero padding of the byte planes when k is not a multiple of 16 (imma_kp): only ever non-empty for toy parameter sets
void planes_clear(pvw_ctx* c, DevBuf& buf, size_t bytes) {
  buf.ensure(bytes);
  if (imma_kp(c->hp.k) != c->hp.k) CUDA_CHECK(cudaMemsetAsync(buf.p, 0, bytes, c->stream));
}
const uint8_t* planes_A(pvw_ctx* c) {
  const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell, kp = imma_kp(k);
  if (!c->As_valid) {
    planes_clear(c, c->As, (size_t)L * ell * k * 8 * kp);
    bool ok = true;
    launch(c, PVW_KERNEL_PERMUTE, 0.0, [&] { ok = launch_imma_planes_m(c->A.as<u64>(), (size_t)k * k * ell, (size_t)k * ell, k, k, L, ell, c->As.as<uint8_t>(), (size_t)k * 8 * kp, true, c->stream); });
    require(ok, PVW_ERR_INTERNAL, "byte-plane conversion of A: launch grid too large");
    c->As_valid = true;
  }
  return c->As.as<uint8_t>();
}
const uint8_t* planes_B(pvw_ctx* c) {
  const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell, nrows = c->nrows, kp = imma_kp(k);
  if (!c->Bs_valid) {
    planes_clear(c, c->Bs, (size_t)L * ell * nrows * 8 * kp);
    bool ok = true;
    launch(c, PVW_KERNEL_PERMUTE, 0.0, [&] { ok = launch_imma_planes_m(c->B.as<u64>(), (size_t)nrows * k * ell, (size_t)k * ell, nrows, k, L, ell, c->Bs.as<uint8_t>(), (size_t)nrows * 8 * kp, true, c->stream); });
    require(ok, PVW_ERR_INTERNAL, "byte-plane conversion of B: launch grid too large");
    c->Bs_valid = true;
  }
  if (c->planes_only && !c->B_released && c->B.p) {   // the batched path reads the planes only: keep ONE resident copy of B
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->B.release();
    c->B_released = true;
  }
  return c->Bs.as<uint8_t>();
}
void imma_launch(pvw_ctx* c, ImmaArgs g) {
  g.pair = c->imma_pair;
  g.stages = c->imma_stages;
  g.epi_warps = c->imma_epi_warps;
  g.fast_reduce = c->imma_fast_reduce && *std::min_element(c->hp.moduli.begin(), c->hp.moduli.end()) >= (1ull << 61);
  bool ok = true;
  launch(c, PVW_KERNEL_IMMA, (double)g.D * g.rows * (g.k + 1.0) * g.L * g.ell * 8.0, [&] { ok = launch_imma_gemm(g, c->stream); });
  require(ok, PVW_ERR_INTERNAL, "tensor-map creation failed for the tensor-core product");
}

void ensure_At(pvw_ctx* c) {
  if (c->At_valid) return;
  const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell;
  const size_t kk = (size_t)k * k * ell;
  c->At.ensure((size_t)L * kk * 8);
  // At[limb][cidx][j] = A[limb][j][cidx]; x = j walks the output's contiguous axis
  launch(c, PVW_KERNEL_PERMUTE, 0.0, [&] { launch_permute(c->A.as<u64>(), c->At.as<u64>(), L, k, k, ell, kk, (size_t)k * ell, ell, kk, ell, (size_t)k * ell, c->stream); });
  c->At_valid = true;
}

template <class F>
int guarded(pvw_ctx* c, F&& f) {
  if (!c) return PVW_ERR_INVALID_PARAMETERS;
  try {
    c->use();
    f();
    return PVW_OK;
  } catch (const PvwException& e) {
    c->err = e.what();
    return e.code;
  } catch (const std::exception& e) {
    c->err = e.what();
    return PVW_ERR_INTERNAL;
  } catch (...) {
    c->err = "unknown failure";
    return PVW_ERR_INTERNAL;
  }
}

}  // namespace

extern "C" {

static void shard_push_c1(pvw_ctx* c, uint32_t slot0, uint32_t count);

int pvw_ctx_create(pvw_ctx** out, const pvw_params_desc* d) {
  if (!out || !d) { g_create_error = "null argument"; return PVW_ERR_INVALID_PARAMETERS; }
  *out = nullptr;
  pvw_ctx* c = nullptr;
  try {
    c = new pvw_ctx();
    c->hp.build(d->n, d->k, d->ell, d->L, d->moduli, d->psi, d->secret_variance, d->error_bound_1, d->error_bound_2);
    c->row0 = d->row0;
    c->nrows = d->nrows == 0 ? d->n - std::min(d->row0, d->n) : d->nrows;
    require((uint64_t)c->row0 + c->nrows <= d->n && c->nrows > 0, PVW_ERR_INVALID_PARAMETERS,
            fmt("party shard [%u, %u) is outside [0, n=%u)", c->row0, c->row0 + c->nrows, d->n));
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw PvwException(PVW_ERR_INTERNAL, std::string("no CUDA device available (there is no CPU fallback): ") +
                                               (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    require(d->device >= 0 && d->device < ndev, PVW_ERR_INVALID_PARAMETERS, fmt("device %d out of range (%d devices)", d->device, ndev));
    c->device = d->device;
    c->use();
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto& e : c->ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    upload_tables(c);
    upload_wire_tables(c);
    *out = c;
    return PVW_OK;
  } catch (const PvwException& e) {
    g_create_error = e.what();
    int code = e.code;
    delete c;
    return code;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    delete c;
    return PVW_ERR_INTERNAL;
  }
}

static void shard_disconnect(pvw_ctx* c);

void pvw_ctx_destroy(pvw_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  shard_disconnect(c);
  if (c->sh.flags) cudaFree(c->sh.flags);
  if (c->sh.xstream) cudaStreamDestroy(c->sh.xstream);
  if (c->sh.ev_c1) cudaEventDestroy(c->sh.ev_c1);
  if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  for (auto& e : c->chunk_ev) cudaEventDestroy(e);
  for (auto& r : c->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
  for (DevBuf* b : {&c->tables, &c->A, &c->At, &c->B, &c->c1s, &c->c2s, &c->stage, &c->rhat, &c->in_small, &c->in_small2, &c->in_m,
                    &c->shat, &c->z, &c->y, &c->X, &c->outd, &c->idxd, &c->idxp, &c->fb, &c->prod1, &c->fscr, &c->narrow, &c->wire_tab, &c->wire_buf, &c->As, &c->Bs, &c->Vx, &c->shat_s, &c->prod})
    b->release();
  delete c;
}

const char* pvw_last_error(const pvw_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int pvw_params_bigint(const pvw_ctx* c, int which, uint64_t* out, uint32_t cap, uint32_t* nwords) {
  if (!c) return PVW_ERR_INVALID_PARAMETERS;
  const BigU* b = which == 0 ? &c->hp.Q : which == 1 ? &c->hp.delta : which == 2 ? &c->hp.delta_pow : nullptr;
  if (!b) return PVW_ERR_INVALID_PARAMETERS;
  if (nwords) *nwords = (uint32_t)b->w.size();
  if (out) {
    if (cap < b->w.size()) return PVW_ERR_DIMENSION_MISMATCH;
    b->to_words(out, cap);
  }
  return PVW_OK;
}
int pvw_params_psi(const pvw_ctx* c, uint64_t* out) {
  if (!c || !out) return PVW_ERR_INVALID_PARAMETERS;
  memcpy(out, c->hp.psi.data(), (size_t)c->hp.L * 8);
  return PVW_OK;
}
int pvw_params_correctness_condition(const pvw_ctx* c, int* ok) {
  if (!c || !ok) return PVW_ERR_INVALID_PARAMETERS;
  *ok = c->hp.correctness_condition() ? 1 : 0;
  return PVW_OK;
}

int pvw_crs_upload(pvw_ctx* c, const uint64_t* A, uint32_t flags) {
  return guarded(c, [&] {
    require(A != nullptr, PVW_ERR_INVALID_PARAMETERS, "A is null");
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell;
    const size_t kk = (size_t)k * k;
    c->A.ensure((size_t)L * kk * ell * 8);
    upload_polys(c, A, kk, c->A.as<u64>(), kk * ell, flags, true);
    c->A_set = true;
    c->At_valid = false; c->As_valid = false;
  });
}
int pvw_crs_download(pvw_ctx* c, uint64_t* A) {
  return guarded(c, [&] {
    require(c->A_set, PVW_ERR_INVALID_PARAMETERS, "CRS has not been uploaded");
    const size_t kk = (size_t)c->hp.k * c->hp.k;
    download_polys(c, c->A.as<u64>(), kk * c->hp.ell, kk, A, true);
  });
}

int pvw_crs_tag_to_seed(const char* tag, uint8_t seed_out[32]) {
  if (!tag || !seed_out) return PVW_ERR_INVALID_PARAMETERS;
  crs_tag_to_seed(tag, seed_out);
  return PVW_OK;
}
int pvw_crs_expand_seed(uint32_t k, uint32_t ell, uint32_t L, const uint64_t* moduli, const uint8_t seed[32], uint64_t* A_out) {
  if (!moduli || !seed || !A_out || k == 0 || ell == 0 || L == 0) return PVW_ERR_INVALID_PARAMETERS;
  for (uint32_t j = 0; j < L; j++) if (moduli[j] < 2) return PVW_ERR_INVALID_PARAMETERS;
  crs_new_deterministic(seed, k, moduli, L, ell, A_out);
  return PVW_OK;
}
int pvw_crs_generate_deterministic(pvw_ctx* c, const uint8_t seed[32], uint64_t* A_out) {
  return guarded(c, [&] {
    require(seed != nullptr, PVW_ERR_INVALID_PARAMETERS, "seed is null");
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell;
    const size_t kk = (size_t)k * k;
    std::vector<uint64_t> local;
    uint64_t* A = A_out;
    if (!A) { local.resize(kk * L * ell); A = local.data(); }
    crs_new_deterministic(seed, k, c->hp.moduli.data(), L, ell, A);
    c->A.ensure((size_t)L * kk * ell * 8);
    upload_polys(c, A, kk, c->A.as<u64>(), kk * ell, PVW_IO_HOST, true);
    c->A_set = true;
    c->At_valid = false; c->As_valid = false;
  });
}
int pvw_crs_generate_from_tag(pvw_ctx* c, const char* tag, uint64_t* A_out) {
  if (!tag) return PVW_ERR_INVALID_PARAMETERS;
  uint8_t seed[32];
  crs_tag_to_seed(tag, seed);
  return pvw_crs_generate_deterministic(c, seed, A_out);
}

static void ensure_B(pvw_ctx* c) {
  const size_t bytes = (size_t)c->hp.L * c->nrows * c->hp.k * c->hp.ell * 8;
  if (c->B_released) {   // a single call, a download or a key update needs the u64 operand again: rebuild it from the planes
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell, nrows = c->nrows, kp = imma_kp(k);
    c->B.ensure(bytes);
    bool ok = true;
    launch(c, PVW_KERNEL_PERMUTE, 0.0, [&] { ok = launch_imma_unplanes_m(c->B.as<u64>(), (size_t)nrows * k * ell, (size_t)k * ell, nrows, k, L, ell, c->Bs.as<uint8_t>(), (size_t)nrows * 8 * kp, true, c->stream); });
    require(ok, PVW_ERR_INTERNAL, "rebuilding B from its byte planes: launch grid too large");
    c->B_released = false;
    return;
  }
  if (c->B.bytes >= bytes) return;
  c->B.ensure(bytes);
  CUDA_CHECK(cudaMemsetAsync(c->B.p, 0, bytes, c->stream));  // GlobalPublicKey::new fills with zero polys, public_key.rs:196-208
  c->Bs_valid = false;
}
static void check_rows(pvw_ctx* c, uint32_t row, uint32_t count) {
  if ((uint64_t)row + count > c->hp.n)  // add_public_key: index >= n, public_key.rs:216-221
    throw PvwException(PVW_ERR_INDEX_OUT_OF_BOUNDS, fmt("party index %u exceeds n=%u", row + count - 1, c->hp.n));
  require(row >= c->row0 && (uint64_t)row + count <= (uint64_t)c->row0 + c->nrows, PVW_ERR_INDEX_OUT_OF_BOUNDS,
          fmt("rows [%u, %u) are outside this context's shard [%u, %u)", row, row + count, c->row0, c->row0 + c->nrows));
}

int pvw_pk_upload_rows(pvw_ctx* c, uint32_t row, uint32_t count, const uint64_t* B, uint32_t flags) {
  return guarded(c, [&] {
    if (count == 0) return;
    require(B != nullptr, PVW_ERR_INVALID_PARAMETERS, "B is null");
    check_rows(c, row, count);
    ensure_B(c);
    const uint32_t k = c->hp.k, ell = c->hp.ell;
    upload_polys(c, B, (uint64_t)count * k, c->B.as<u64>() + (size_t)(row - c->row0) * k * ell, (size_t)c->nrows * k * ell, flags, true);
    c->num_keys = std::max(c->num_keys, row + count); c->Bs_valid = false;
  });
}
int pvw_pk_download_rows(pvw_ctx* c, uint32_t row, uint32_t count, uint64_t* B) {
  return guarded(c, [&] {
    if (count == 0) return;
    check_rows(c, row, count);
    ensure_B(c);
    const uint32_t k = c->hp.k, ell = c->hp.ell;
    download_polys(c, c->B.as<u64>() + (size_t)(row - c->row0) * k * ell, (size_t)c->nrows * k * ell, (uint64_t)count * k, B, true);
  });
}
int pvw_pk_num_keys(const pvw_ctx* c, uint32_t* num_keys) {
  if (!c || !num_keys) return PVW_ERR_INVALID_PARAMETERS;
  *num_keys = c->num_keys;
  return PVW_OK;
}

int pvw_keygen_batch(pvw_ctx* c, uint32_t row, uint32_t count, const void* sk, const void* e, uint32_t flags) {
  return guarded(c, [&] {
    if (count == 0) return;
    require(sk && e, PVW_ERR_INVALID_PARAMETERS, "sk / e is null");
    require(c->A_set, PVW_ERR_KEYGEN, "CRS has not been uploaded");
    check_rows(c, row, count);
    ensure_B(c);
    ensure_At(c);
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell;
    const int sb = secret_bytes(flags), eb = error_bytes(flags);
    const size_t small = (size_t)count * k * ell;
    const void* d_sk = stage_in(c, c->in_small, sk, small * sb, flags);
    const void* d_e = stage_in(c, c->in_small2, e, small * eb, flags);
    // s_hat[p][limb][j][ell]
    c->rhat.ensure((size_t)count * L * k * ell * 8);
    ntt(c, d_sk, sb, nullptr, (uint64_t)count * k, k, c->rhat.as<u64>(), (size_t)L * k * ell, (size_t)k * ell, false, true);
    // B rows <- NTT(e): item idx = p*k + cidx lands at B[limb][row+p][cidx]
    u64* Brow = c->B.as<u64>() + (size_t)(row - c->row0) * k * ell;
    ntt(c, d_e, eb, nullptr, (uint64_t)count * k, (uint32_t)std::min<uint64_t>((uint64_t)count * k, 0xFFFFFFFFu), Brow, 0, (size_t)c->nrows * k * ell);
    GemmArgs g{};
    g.M = c->At.as<u64>(); g.M_ls = (size_t)k * k * ell; g.M_rs = (size_t)k * ell;
    g.V = c->rhat.as<u64>(); g.V_ls = (size_t)k * ell; g.V_ds = (size_t)L * k * ell;
    g.O = Brow; g.O_ls = (size_t)c->nrows * k * ell; g.O_ds = (size_t)k * ell; g.O_packed = 1;  // B is an operand of the c2 product
    g.rows = k; g.D = count; g.k = k; g.L = L; g.ell = ell; g.mode = 0; g.lc = c->T.lc;
    gemm(c, g);
    c->num_keys = std::max(c->num_keys, row + count); c->Bs_valid = false;
    if (!(flags & PVW_IO_DEVICE)) CUDA_CHECK(cudaStreamSynchronize(c->stream));  // host buffers may be reused by the caller
  });
}

int pvw_crs_multiply_by_randomness(pvw_ctx* c, uint32_t D, const uint64_t* r_hat, uint64_t* out) {
  return guarded(c, [&] {
    if (D == 0) return;
    require(r_hat && out, PVW_ERR_INVALID_PARAMETERS, "null argument");
    require(c->A_set, PVW_ERR_INVALID_PARAMETERS, "CRS has not been uploaded");
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell;
    const size_t per = (size_t)k * L * ell;  // words per dealer
    c->stage.ensure((size_t)D * per * 8);
    c->rhat.ensure((size_t)D * per * 8);
    c->z.ensure((size_t)D * per * 8);
    CUDA_CHECK(cudaMemcpyAsync(c->stage.p, r_hat, (size_t)D * per * 8, cudaMemcpyHostToDevice, c->stream));
    // [D][k][L][ell] -> [D][L][k][ell]
    launch(c, PVW_KERNEL_PERMUTE, 0.0, [&] { launch_permute(c->stage.as<u64>(), c->rhat.as<u64>(), D, k, L, ell, per, (size_t)L * ell, ell, per, ell, (size_t)k * ell, c->stream, 1); });
    CUDA_CHECK(cudaMemsetAsync(c->z.p, 0, (size_t)D * per * 8, c->stream));
    GemmArgs g{};
    g.M = c->A.as<u64>(); g.M_ls = (size_t)k * k * ell; g.M_rs = (size_t)k * ell;
    g.V = c->rhat.as<u64>(); g.V_ls = (size_t)k * ell; g.V_ds = per;
    g.O = c->z.as<u64>(); g.O_ls = (size_t)k * ell; g.O_ds = per;
    g.rows = k; g.D = D; g.k = k; g.L = L; g.ell = ell; g.mode = 0; g.lc = c->T.lc;
    gemm(c, g);
    launch(c, PVW_KERNEL_PERMUTE, 0.0, [&] { launch_permute(c->z.as<u64>(), c->stage.as<u64>(), D, L, k, ell, per, (size_t)k * ell, ell, per, ell, (size_t)L * ell, c->stream); });
    CUDA_CHECK(cudaMemcpyAsync(out, c->stage.p, (size_t)D * per * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
}

int pvw_ct_reserve(pvw_ctx* c, uint32_t capacity) {
  return guarded(c, [&] {
    const size_t w1 = c->c1_stride(), w2 = (size_t)c->hp.L * c->nrows * c->hp.ell;
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    require(!c->sh.connected, PVW_ERR_INVALID_PARAMETERS, "pvw_ct_reserve: peers hold IPC mappings of this store; call pvw_shard_disconnect on every rank first");
    c->c1s.release(); c->c2s.release();
    c->cap = 0; c->c1p_valid.clear();
    if (capacity == 0) return;
    c->c1s.ensure((size_t)capacity * w1 * 8);
    c->c2s.ensure((size_t)capacity * w2 * 8);
    CUDA_CHECK(cudaMemsetAsync(c->c1s.p, 0, (size_t)capacity * w1 * 8, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->c2s.p, 0, (size_t)capacity * w2 * 8, c->stream));
    c->cap = capacity;
    c->c1p_valid.assign(capacity, 0);
  });
}

int pvw_encrypt_batch(pvw_ctx* c, uint32_t slot0, uint32_t D, uint32_t c1_lo, uint32_t c1_hi, const uint64_t* m, const void* r,
                      const void* e1, const void* e2, uint32_t flags) {
  return guarded(c, [&] {
    if (D == 0) return;
    const HostParams& hp = c->hp;
    const uint32_t L = hp.L, k = hp.k, ell = hp.ell, nrows = c->nrows;
    require(!((flags & PVW_ENC_C1_ONLY) && (flags & PVW_ENC_C2_ONLY)), PVW_ERR_INVALID_PARAMETERS, "C1_ONLY and C2_ONLY exclude each other");
    require(!(flags & PVW_ENC_PUSH_C1) || (c->sh.connected && !(flags & PVW_ENC_C2_ONLY)), PVW_ERR_INVALID_PARAMETERS,
            "PVW_ENC_PUSH_C1 needs a connected shard exchange (pvw_shard_connect) and a call that computes c1");
    require(c1_lo <= c1_hi && c1_hi <= D, PVW_ERR_INVALID_PARAMETERS, "bad c1 dealer range");
    const bool do_c2 = !(flags & PVW_ENC_C1_ONLY);
    if (flags & PVW_ENC_C2_ONLY) c1_lo = c1_hi = 0;
    require(r && (!do_c2 || (m && e2)), PVW_ERR_INVALID_PARAMETERS, "null argument");
    require(c1_lo == c1_hi || e1 != nullptr, PVW_ERR_INVALID_PARAMETERS, "e1 is null");
    require(c->A_set, PVW_ERR_INVALID_PARAMETERS, "CRS has not been uploaded");
    // encryption.rs:117-121: is_full() <=> num_keys >= n ; restricted to this shard: every local row below num_keys
    require(c->num_keys >= c->row0 + nrows, PVW_ERR_INVALID_PARAMETERS, "Global public key is not complete (missing party keys)");
    // encryption.rs:124-128
    require(hp.correctness_condition(), PVW_ERR_INVALID_PARAMETERS,
            "Parameters do not satisfy correctness condition - decryption may fail");
    require((uint64_t)slot0 + D <= c->cap, PVW_ERR_INDEX_OUT_OF_BOUNDS,
            fmt("ciphertext slots [%u, %u) exceed the reserved capacity %u", slot0, slot0 + D, c->cap));
    const size_t w1 = (size_t)L * k * ell, w2 = (size_t)L * nrows * ell, s1 = c->c1_stride();
    const bool host = !(flags & PVW_IO_DEVICE);
    // copy order matters (one host-to-device DMA queue): the small inputs the first kernels need go first, then the big
    // ones (e2, m) on the copy stream, waited for only after the c2 product
    const int sb = secret_bytes(flags), eb = error_bytes(flags);
    // the staging buffers may still feed kernels of an earlier host-pointer encryption (which no longer waits for them, see the end)
    if (host && c->tail_pending) CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev[2], 0));
    const void* d_r = stage_in(c, c->in_small, r, (size_t)D * k * ell * sb, flags);
    const void* d_e1 = nullptr;
    if (c1_hi > c1_lo)
      d_e1 = stage_in(c, c->in_small2, at_bytes(e1, (size_t)c1_lo * k * ell * eb), (size_t)(c1_hi - c1_lo) * k * ell * eb, flags);
    if (host) CUDA_CHECK(cudaEventRecord(c->ev[1], c->stream));   // r and e1 have left the caller's buffers once this has fired
    const void* d_e2 = do_c2 ? stage_in_overlapped(c, c->stage, e2, (size_t)D * nrows * ell * eb, flags, c->ev[0]) : nullptr;
    const u64* d_m = do_c2 ? (const u64*)stage_in_overlapped(c, c->in_m, m, (size_t)D * nrows * 8, flags, c->ev[0]) : nullptr;
    // r_hat (encryption.rs:147-154): operand form [d][limb][j][ell] for the IMAD kernel, byte planes for the tensor-core one
    const bool imma = imma_wanted(c, nrows, D);
    const uint32_t kp = imma_kp(k);
    // c1 alone (PVW_ENC_C1_ONLY, the first half of a multi-GPU step): only the slice's dealers need r_hat
    const bool slice_only = imma && !do_c2;
    const uint32_t v_first = slice_only ? c1_lo : 0, v_count = slice_only ? c1_hi - c1_lo : D;
    if (imma) {
      planes_clear(c, c->Vx, (size_t)L * ell * v_count * 8 * kp);
      if (v_count)
        ntt(c, at_bytes(d_r, (size_t)v_first * k * ell * sb), sb, nullptr, (uint64_t)v_count * k, k, c->Vx.as<u64>(), kp, (size_t)v_count * 8 * kp, false, false, 2);
    } else {
      c->rhat.ensure((size_t)D * w1 * 8);
      ntt(c, d_r, sb, nullptr, (uint64_t)D * k, k, c->rhat.as<u64>(), w1, (size_t)k * ell, false, true);
    }
    u64* c1 = c->c1s.as<u64>() + (size_t)slot0 * s1;
    u64* c2 = c->c2s.as<u64>() + (size_t)slot0 * w2;
    // tensor-core path, usual shapes: the c1 product goes to a slot-major scratch and ONE kernel finishes c1 = NTT(e1) + product
    // into the store, as residues and as byte planes (ntt_c1_finish_kernel).  Otherwise: c1 <- NTT(e1) (encryption.rs:161-167)
    // here and c1 += A r_hat in the product's epilogue (crs.rs:187-199, encryption.rs:171-173).
    const bool c1_direct = imma && c1_hi > c1_lo && k % 4 == 0 && ell <= 16;
    if (c1_hi > c1_lo && !c1_direct)
      ntt(c, d_e1, eb, nullptr, (uint64_t)(c1_hi - c1_lo) * k, k, c1 + (size_t)c1_lo * s1, s1, (size_t)k * ell);
    if (c1_hi > c1_lo && !c1_direct) c->c1_planes_invalidate(slot0 + c1_lo, c1_hi - c1_lo);
    // device inputs: c2 <- NTT(e2) + (m as i64) * g_hat   (encryption.rs:195-196), then c2 += B r_hat   (:185-192, :198)
    // host inputs:   c2 <- B r_hat first (it does not need e2 / m, whose copy is still in flight), then c2 += NTT(e2) + m g_hat
    auto preload = [&](bool accumulate) {
      ntt(c, d_e2, eb, d_m, (uint64_t)D * nrows, nrows, c2, w2, (size_t)nrows * ell, accumulate);
    };
    if (!host && !imma && do_c2) preload(false);
    if (imma) {
      // tensor-core product (imma.cu) on the byte planes of A / B (built once) and of r_hat (written by the NTT kernel)
      ImmaArgs g{};
      g.Vb = c->Vx.as<uint8_t>(); g.Vb_plane = (size_t)v_count * 8 * kp; g.Vb_D = v_count;
      g.k = k; g.L = L; g.ell = ell; g.lc = c->T.lc;
      if (c1_hi > c1_lo) {
        g.Mb = planes_A(c); g.Mb_plane = (size_t)k * 8 * kp; g.rows = k;
        g.d_first = c1_lo - v_first; g.D = c1_hi - c1_lo;
        if (c1_direct) {
          const uint32_t Dc1 = c1_hi - c1_lo;
          c->prod1.ensure((size_t)Dc1 * w1 * 8);
          g.O = c->prod1.as<u64>(); g.O_ls = (size_t)ell * k; g.O_ds = (size_t)L * ell * k; g.O_rs = 1; g.O_cs = k; g.O_packed = 0; g.mode = 2;
          imma_launch(c, g);
          bool ok = true;
          launch(c, PVW_KERNEL_NTT, 0.0, [&] { ok = launch_ntt_c1_finish(c->T, d_e1, eb, (uint64_t)Dc1 * k, k, c1 + (size_t)c1_lo * s1, s1, c->prod1.as<u64>(), kp, c->stream); });
          require(ok, PVW_ERR_INTERNAL, "c1 finisher: unsupported shape");
          for (uint32_t i = slot0 + c1_lo; i < slot0 + c1_hi; i++) c->c1p_valid[i] = 1;
        } else {
          g.O = c1 + (size_t)c1_lo * s1; g.O_ls = (size_t)k * ell; g.O_ds = s1; g.O_rs = ell; g.O_cs = 1; g.O_packed = 1; g.mode = 0;
          imma_launch(c, g);
        }
      }
      if (flags & PVW_ENC_PUSH_C1) {   // peer copies start now, under the c2 product
        shard_push_c1(c, slot0 + c1_lo, c1_hi - c1_lo);
        c->sh.batch_slot0 = slot0; c->sh.batch_D = D; c->sh.batch_direct = c1_direct;
      }
      if (do_c2) {
        // The product goes to a slot-major scratch (lanes of a warp = consecutive parties: full-sector stores), a chunk of
        // dealers at a time; the NTT kernel then writes c2 = NTT(e2) + m g_hat + product in the store layout, reading the product
        // with unit stride.  (Storing 8-byte results straight into the store layout, 64 bytes apart, cost 10 GB of DRAM traffic
        // per launch in partial-sector writes and fills and made the launch DRAM-bound.)
        g.Mb = planes_B(c); g.Mb_plane = (size_t)nrows * 8 * kp; g.rows = nrows;
        g.O_ls = (size_t)ell * nrows; g.O_ds = (size_t)L * ell * nrows; g.O_rs = 1; g.O_cs = nrows; g.O_packed = 0; g.mode = 2;
        const uint32_t step = (uint32_t)std::min<int64_t>(D, std::max<int64_t>(16, c->imma_chunk_dealers));
        c->prod.ensure((size_t)step * w2 * 8);
        for (uint32_t dc0 = 0; dc0 < D; dc0 += step) {
          const uint32_t Dc = std::min(step, D - dc0);
          g.d_first = dc0; g.D = Dc; g.O = c->prod.as<u64>();
          imma_launch(c, g);
          if (host && dc0 == 0) CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev[0], 0));   // e2 / m arrive under the first product
          ntt(c, at_bytes(d_e2, (size_t)dc0 * nrows * ell * eb), eb, d_m + (size_t)dc0 * nrows, (uint64_t)Dc * nrows, nrows, c2 + (size_t)dc0 * w2, w2,
              (size_t)nrows * ell, false, false, 0, c->prod.as<u64>());
        }
      }
    } else {
      if (c1_hi > c1_lo) {
        const uint32_t Dc = c1_hi - c1_lo;
        GemmArgs g{};
        g.M = c->A.as<u64>(); g.M_ls = (size_t)k * k * ell; g.M_rs = (size_t)k * ell;
        g.V = c->rhat.as<u64>() + (size_t)c1_lo * w1; g.V_ls = (size_t)k * ell; g.V_ds = w1;
        g.O = c1 + (size_t)c1_lo * s1; g.O_ls = (size_t)k * ell; g.O_ds = s1; g.O_packed = 1;  // c1 is an operand of the decrypt product
        g.rows = k; g.D = Dc; g.k = k; g.L = L; g.ell = ell; g.mode = 0; g.lc = c->T.lc;
        gemm(c, g);
      }
      if (flags & PVW_ENC_PUSH_C1) {
        shard_push_c1(c, slot0 + c1_lo, c1_hi - c1_lo);
        c->sh.batch_slot0 = slot0; c->sh.batch_D = D; c->sh.batch_direct = false;
      }
      if (do_c2) {
        ensure_B(c);
        GemmArgs g{};
        g.M = c->B.as<u64>(); g.M_ls = (size_t)nrows * k * ell; g.M_rs = (size_t)k * ell;
        g.V = c->rhat.as<u64>(); g.V_ls = (size_t)k * ell; g.V_ds = w1;
        g.O = c2; g.O_ls = (size_t)nrows * ell; g.O_ds = w2;
        g.rows = nrows; g.D = D; g.k = k; g.L = L; g.ell = ell; g.mode = host ? 2 : 0; g.lc = c->T.lc;
        gemm(c, g);
      }
    }
    if (host && do_c2 && !imma) {
      CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev[0], 0));
      preload(true);
    }
    if (host) {
      // The call returns once the caller's buffers have been read -- r, e1 (head of the compute stream) and e2, m (copy stream) --
      // not once the kernels have run: the ciphertexts stay in the device store, every consumer is ordered on the same stream, and
      // the host side of the next call (a decryption's index and key copies) is then queued under this call's products.  The
      // staging buffers are protected by ev[2]: the copy stream of the next host-pointer call waits for it.
      CUDA_CHECK(cudaEventRecord(c->ev[2], c->stream));
      c->tail_pending = true;
      CUDA_CHECK(cudaEventSynchronize(c->ev[1]));
      if (do_c2) CUDA_CHECK(cudaEventSynchronize(c->ev[0]));
    }
  });
}

int pvw_ct_download(pvw_ctx* c, uint32_t slot, uint64_t* c1, uint64_t* c2) {
  return guarded(c, [&] {
    require(slot < c->cap, PVW_ERR_INDEX_OUT_OF_BOUNDS, fmt("ciphertext slot %u exceeds the reserved capacity %u", slot, c->cap));
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell, nrows = c->nrows;
    if (c1) download_polys(c, c->c1s.as<u64>() + (size_t)slot * c->c1_stride(), (size_t)k * ell, k, c1, true);
    if (c2) download_polys(c, c->c2s.as<u64>() + (size_t)slot * L * nrows * ell, (size_t)nrows * ell, nrows, c2, false);
  });
}
int pvw_ct_upload(pvw_ctx* c, uint32_t slot, const uint64_t* c1, const uint64_t* c2) {
  return guarded(c, [&] {
    require(slot < c->cap, PVW_ERR_INDEX_OUT_OF_BOUNDS, fmt("ciphertext slot %u exceeds the reserved capacity %u", slot, c->cap));
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell, nrows = c->nrows;
    if (c1) { upload_polys(c, c1, k, c->c1s.as<u64>() + (size_t)slot * c->c1_stride(), (size_t)k * ell, PVW_IO_HOST, true); c->c1_planes_invalidate(slot, 1); }
    if (c2) upload_polys(c, c2, nrows, c->c2s.as<u64>() + (size_t)slot * L * nrows * ell, (size_t)nrows * ell, PVW_IO_HOST, false);
  });
}
int pvw_ct_c1_device_ptr(pvw_ctx* c, uint32_t slot, void** ptr, uint64_t* slot_stride) {
  return guarded(c, [&] {
    require(ptr != nullptr, PVW_ERR_INVALID_PARAMETERS, "null argument");
    require(slot < c->cap, PVW_ERR_INDEX_OUT_OF_BOUNDS, fmt("ciphertext slot %u exceeds the reserved capacity %u", slot, c->cap));
    const size_t w1 = c->c1_stride();
    *ptr = c->c1s.as<u64>() + (size_t)slot * w1;
    if (slot_stride) *slot_stride = w1;
    c->c1_external = true;   // the caller may now write c1 residues (an NCCL all-gather): the planes behind them are no longer trusted
  });
}

static void decode_on_device(pvw_ctx* c, const u64* z, size_t z_ls, size_t z_ds, uint32_t Pc, uint32_t D, u64* out, size_t out_ps,
                             size_t z_cs = 0, const DecodeSub* sub = nullptr) {
  const uint64_t S = (uint64_t)Pc * D;
  c->y.ensure(decode_scratch_words_y(c->T, S) * 8);
  c->X.ensure(decode_scratch_words_X(c->T, S) * 8);
  // fast path first: one kernel decodes every share a correct run produces and lists the others; the chain then runs on the list
  FallbackList fb{nullptr, nullptr};
  if (c->F.enabled && c->decode_fused && S < (1ull << 32)) {
    c->fb.ensure(16 + S * 4);
    uint32_t* count = c->fb.as<uint32_t>();
    uint32_t* list = count + 4;
    CUDA_CHECK(cudaMemsetAsync(count, 0, 4, c->stream));
    bool fused = false;
    u64* scr = nullptr;
    if (c->T.ell == 8 || c->T.ell == 16) {   // thread-per-share form: two launches with l + 2 scratch words per share between them
      c->fscr.ensure(decode_fused_scratch_words(c->T, S) * 8);
      scr = c->fscr.as<u64>();
    }
    launch(c, PVW_KERNEL_DECODE_FUSED, 0.0, [&] { fused = launch_decode_fused(c->T, c->F, z, z_ls, z_ds, Pc, D, out, out_ps, list, count, c->stream, z_cs, sub, scr); });
    if (fused && scr) c->launches++;   // the two-launch form queued a second kernel
    if (fused) fb = FallbackList{list, count};
  }
  const FallbackList* fbp = fb.count ? &fb : nullptr;
  bool ok = true;
  launch(c, PVW_KERNEL_DECODE_RNS, 0.0, [&] { ok = launch_decode_rns(c->T, z, z_ls, z_ds, Pc, D, c->y.as<u64>(), c->stream, z_cs, sub, fbp); });
  require(ok, PVW_ERR_INTERNAL, "decode: unsupported shape");
  launch(c, PVW_KERNEL_CRT_LIFT, 0.0, [&] { launch_crt_lift(c->T, c->y.as<u64>(), c->X.as<u64>(), S, c->stream, fbp); });
  launch(c, PVW_KERNEL_DECODE_TAIL, 0.0, [&] { launch_decode_tail(c->T, c->X.as<u64>(), Pc, D, out, out_ps, c->stream, fbp); });
}

int pvw_decrypt_batch(pvw_ctx* c, uint32_t D, const uint32_t* dealer_slots, uint32_t P, const uint32_t* party_idx, const void* sk,
                      uint64_t* out, uint32_t flags) {
  return guarded(c, [&] {
    if (D == 0 || P == 0) return;
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell, nrows = c->nrows;
    const size_t s1 = c->c1_stride();
    require(party_idx && sk && out, PVW_ERR_INVALID_PARAMETERS, "null argument");
    if (dealer_slots) {
      for (uint32_t d = 0; d < D; d++)
        require(dealer_slots[d] < c->cap, PVW_ERR_INDEX_OUT_OF_BOUNDS, fmt("ciphertext slot %u exceeds the reserved capacity %u", dealer_slots[d], c->cap));
    } else {
      require(D <= c->cap, PVW_ERR_INDEX_OUT_OF_BOUNDS, fmt("%u ciphertexts requested, %u reserved", D, c->cap));
    }
    std::vector<uint32_t>& local = c->h_idxp;   // owned by the context: nothing here goes out of scope under an async copy
    local.resize(P);
    for (uint32_t p = 0; p < P; p++) {
      // decrypt_party_shares: party_index >= n, decryption.rs:303-309
      if (party_idx[p] >= c->hp.n) throw PvwException(PVW_ERR_INVALID_PARAMETERS, fmt("Party index %u exceeds maximum %u", party_idx[p], c->hp.n - 1));
      require(party_idx[p] >= c->row0 && party_idx[p] < c->row0 + nrows, PVW_ERR_INDEX_OUT_OF_BOUNDS,
              fmt("party %u is outside this context's shard [%u, %u)", party_idx[p], c->row0, c->row0 + nrows));
      local[p] = party_idx[p] - c->row0;
    }
    c->idxp.ensure((size_t)P * 4);
    CUDA_CHECK(cudaMemcpyAsync(c->idxp.p, local.data(), (size_t)P * 4, cudaMemcpyHostToDevice, c->stream));
    const uint32_t* d_slots = nullptr;
    if (dealer_slots) {
      c->idxd.ensure((size_t)D * 4);
      CUDA_CHECK(cudaMemcpyAsync(c->idxd.p, dealer_slots, (size_t)D * 4, cudaMemcpyHostToDevice, c->stream));
      d_slots = c->idxd.as<uint32_t>();
    }
    // (both copies read pageable host memory: cudaMemcpyAsync returns once the source has been staged, so neither the caller's
    //  dealer_slots nor h_idxp needs to outlive this call and the stream is not synchronised)
    const bool host = !(flags & PVW_IO_DEVICE);
    const int sb = secret_bytes(flags);
    const void* d_sk = sk;
    u64* d_out = reinterpret_cast<u64*>(out);
    if (host) {
      c->in_small.ensure((size_t)P * k * ell * sb);
      d_sk = c->in_small.p;
      c->outd.ensure((size_t)P * D * 8);
      d_out = c->outd.as<u64>();
    }
    // tensor-core product: dealers are processed in chunks (their c1 is expanded once per chunk, imma.cu), parties in chunks
    // inside; the IMAD kernel takes all dealers at once
    const bool imma = imma_wanted(c, P, D);
    const uint32_t Dstep = imma ? (uint32_t)std::min<int64_t>(D, std::max<int64_t>(16, c->imma_chunk_dealers)) : D;
    const uint32_t kp = imma_kp(k);
    uint32_t Pc_max = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(P, c->decrypt_chunk_shares / std::max<uint32_t>(Dstep, 1)));
    if (imma) planes_clear(c, c->shat_s, (size_t)L * ell * Pc_max * 8 * kp); else c->shat.ensure((size_t)L * Pc_max * k * ell * 8);
    c->z.ensure((size_t)Dstep * L * Pc_max * ell * 8);
    c->y.ensure(decode_scratch_words_y(c->T, (uint64_t)Pc_max * Dstep) * 8);   // sized for the largest chunk up front: growing a
    c->X.ensure(decode_scratch_words_X(c->T, (uint64_t)Pc_max * Dstep) * 8);   // buffer mid-call would synchronise the device
    c->fb.ensure(16 + (size_t)Pc_max * Dstep * 4);
    c->fscr.ensure(decode_fused_scratch_words(c->T, (uint64_t)Pc_max * Dstep) * 8);
    // host inputs: a short first chunk, so that little of the secret-key copy is exposed before the kernels start, then chunks
    // growing threefold: a party's share of the work takes about 3.5x as long as the copy of its key, so chunk i+1 (<= 3x chunk i)
    // has arrived by the time chunk i is done
    const uint32_t first = host ? std::max<uint32_t>(1, std::min<uint32_t>(Pc_max, std::max<uint32_t>(Pc_max / 8, 64))) : Pc_max;
    std::vector<uint32_t> chunks;
    for (uint32_t p0 = 0, pc = first; p0 < P; ) {
      const uint32_t Pc = std::min(pc, P - p0);
      chunks.push_back(Pc);
      p0 += Pc;
      pc = (uint32_t)std::min<uint64_t>(Pc_max, 3ull * pc);
    }
    if (host) {  // every chunk's secret keys are queued now, in order, each with its own event: chunk i+1 arrives while chunk i computes
      if (c->tail_pending) CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev[2], 0));   // in_small may still hold an encryption's r
      uint32_t i = 0;
      for (uint32_t p0 = 0; p0 < P; i++) {
        const uint32_t Pc = chunks[i];
        if (c->chunk_ev.size() <= i) { cudaEvent_t e; CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->chunk_ev.push_back(e); }
        CUDA_CHECK(cudaMemcpyAsync(c->in_small.as<uint8_t>() + (size_t)p0 * k * ell * sb, at_bytes(sk, (size_t)p0 * k * ell * sb), (size_t)Pc * k * ell * sb,
                                   cudaMemcpyHostToDevice, c->copy_stream));
        CUDA_CHECK(cudaEventRecord(c->chunk_ev[i], c->copy_stream));
        p0 += Pc;
      }
    }
    for (uint32_t dc0 = 0; dc0 < D; dc0 += Dstep) {
      const uint32_t Dc = std::min(Dstep, D - dc0);
      // byte planes of this chunk's c1: read in place when the slots are consecutive and their planes are current (written by the
      // c1 finisher or pushed by a peer), else converted from the residues into a dense buffer
      bool in_place = imma && !d_slots && !c->c1_external;
      for (uint32_t d = dc0; in_place && d < dc0 + Dc; d++) in_place = c->c1p_valid[d] != 0;
      if (imma && !in_place) {
        planes_clear(c, c->Vx, (size_t)L * ell * Dc * 8 * kp);
        bool ok = true;
        launch(c, PVW_KERNEL_EXPAND, (double)Dc * L * k * ell * 16.0, [&] {
          ok = launch_imma_planes_v(d_slots ? c->c1s.as<u64>() : c->c1s.as<u64>() + (size_t)dc0 * s1, s1, (size_t)k * ell, ell, Dc, k, L,
                               c->Vx.as<uint8_t>(), (size_t)Dc * 8 * kp, true, d_slots ? d_slots + dc0 : nullptr, c->stream);
        });
        require(ok, PVW_ERR_INTERNAL, "byte-plane conversion of c1: launch grid too large");
      }
      uint32_t chunk_no = 0;
      for (uint32_t p0 = 0; p0 < P; chunk_no++) {
        const uint32_t Pc = chunks[chunk_no];
        if (host && dc0 == 0) CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->chunk_ev[chunk_no], 0));
        // z[d][limb][p][ell] = sum_j s_hat[p][j] * c1_d[j] - c2_d[party]   (decryption.rs:257-274), with
        // s_hat = SecretKey::get_polynomial (secret_key.rs:98-112) computed once per party, not per ciphertext
        if (imma) {
          const void* sk_chunk = at_bytes(d_sk, (size_t)p0 * k * ell * sb);
          if (sb == 8 && c->narrow_inputs && ntt_planes_take_narrow(c->T, k) && chunk_no < 64) {
            // 64-bit secrets (the reference's i64): the L limbs of the transform would each fetch 8 bytes per coefficient; a one-byte
            // copy is made first (and ignored if a coefficient does not fit)
            c->narrow.ensure((size_t)Pc_max * k * ell + 256);
            uint32_t* nflag = reinterpret_cast<uint32_t*>(c->narrow.as<uint8_t>() + (size_t)Pc_max * k * ell) + chunk_no;
            if (chunk_no == 0) CUDA_CHECK(cudaMemsetAsync(c->narrow.as<uint8_t>() + (size_t)Pc_max * k * ell, 0, 256, c->stream));
            bool ok = true;
            launch(c, PVW_KERNEL_NTT_PLANES, 0.0, [&] { ok = launch_narrow_i64(sk_chunk, c->narrow.p, (uint64_t)Pc * k * ell, nflag, c->stream); });
            require(ok, PVW_ERR_INTERNAL, "narrowing of 64-bit secrets: unsupported shape");
            ntt(c, c->narrow.p, 1, nullptr, (uint64_t)Pc * k, k, c->shat_s.as<u64>(), kp, (size_t)Pc * 8 * kp, false, false, 1, nullptr, sk_chunk, nflag);
          } else {
            ntt(c, sk_chunk, sb, nullptr, (uint64_t)Pc * k, k, c->shat_s.as<u64>(), kp, (size_t)Pc * 8 * kp, false, false, 1);
          }
          // the product is stored alone, slot-major (lanes of a warp = consecutive parties: full-sector stores); c2 is
          // subtracted by the decode kernel, which reads both with unit stride
          ImmaArgs g{};
          g.Mb = c->shat_s.as<uint8_t>(); g.Mb_plane = (size_t)Pc * 8 * kp; g.rows = Pc;
          if (in_place) {
            g.Vb = reinterpret_cast<const uint8_t*>(c->c1s.as<u64>() + (size_t)dc0 * s1 + (size_t)L * k * ell);
            g.Vb_plane = (size_t)8 * kp; g.Vb_dstride = s1 * 8;
          } else {
            g.Vb = c->Vx.as<uint8_t>(); g.Vb_plane = (size_t)Dc * 8 * kp;
          }
          g.Vb_D = Dc; g.d_first = 0; g.D = Dc;
          g.O = c->z.as<u64>(); g.O_ls = (size_t)ell * Pc; g.O_ds = (size_t)L * ell * Pc; g.O_rs = 1; g.O_cs = Pc;
          g.k = k; g.L = L; g.ell = ell; g.mode = 2; g.lc = c->T.lc;
          imma_launch(c, g);
        } else {
          ntt(c, at_bytes(d_sk, (size_t)p0 * k * ell * sb), sb, nullptr, (uint64_t)Pc * k, Pc * k, c->shat.as<u64>(), 0, (size_t)Pc * k * ell, false, true);
          GemmArgs g{};
          g.M = c->shat.as<u64>(); g.M_ls = (size_t)Pc * k * ell; g.M_rs = (size_t)k * ell;
          g.V = c->c1s.as<u64>(); g.V_ls = (size_t)k * ell; g.V_ds = s1; g.V_dmap = d_slots;
          g.O = c->z.as<u64>(); g.O_ls = (size_t)Pc * ell; g.O_ds = (size_t)L * Pc * ell;
          g.S = c->c2s.as<u64>(); g.S_ls = (size_t)nrows * ell; g.S_ds = (size_t)L * nrows * ell; g.S_rowmap = c->idxp.as<uint32_t>() + p0;
          g.rows = Pc; g.D = D; g.k = k; g.L = L; g.ell = ell; g.mode = 1; g.lc = c->T.lc;
          gemm(c, g);
        }
        if (imma) {
          DecodeSub sub{c->c2s.as<u64>(), (size_t)nrows * ell, (size_t)L * nrows * ell, c->idxp.as<uint32_t>() + p0, d_slots ? d_slots + dc0 : nullptr};
          if (!d_slots) sub.S += (size_t)dc0 * sub.S_ds;
          decode_on_device(c, c->z.as<u64>(), (size_t)Pc * ell, (size_t)L * Pc * ell, Pc, Dc, d_out + (size_t)p0 * D + dc0, D, Pc, &sub);
        } else {
          decode_on_device(c, c->z.as<u64>(), (size_t)Pc * ell, (size_t)L * Pc * ell, Pc, Dc, d_out + (size_t)p0 * D + dc0, D);
        }
        if (host) {  // this chunk's plaintexts go home while the next chunk computes
          CUDA_CHECK(cudaEventRecord(c->ev[3], c->stream));
          CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev[3], 0));
          CUDA_CHECK(cudaMemcpy2DAsync(out + (size_t)p0 * D + dc0, (size_t)D * 8, d_out + (size_t)p0 * D + dc0, (size_t)D * 8, (size_t)Dc * 8, Pc,
                                       cudaMemcpyDeviceToHost, c->copy_stream));
        }
        p0 += Pc;
      }
    }
    if (host) {
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->copy_stream));
    }
  });
}

int pvw_decode_batch(pvw_ctx* c, uint32_t count, const uint64_t* zhat, uint64_t* out) {
  return guarded(c, [&] {
    if (count == 0) return;
    require(zhat && out, PVW_ERR_INVALID_PARAMETERS, "null argument");
    const size_t poly = c->poly();
    c->z.ensure((size_t)count * poly * 8);
    c->outd.ensure((size_t)count * 8);
    CUDA_CHECK(cudaMemcpyAsync(c->z.p, zhat, (size_t)count * poly * 8, cudaMemcpyHostToDevice, c->stream));
    // host layout [count][L][ell] read in place: "dealer" = item, one party per dealer
    decode_on_device(c, c->z.as<u64>(), c->hp.ell, poly, 1, count, c->outd.as<u64>(), 0);
    CUDA_CHECK(cudaMemcpyAsync(out, c->outd.p, (size_t)count * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
}

int pvw_ntt_forward_small(pvw_ctx* c, uint32_t count, const int64_t* coeffs, uint64_t* out) {
  return guarded(c, [&] {
    if (count == 0) return;
    require(coeffs && out, PVW_ERR_INVALID_PARAMETERS, "null argument");
    const size_t poly = c->poly();
    c->in_small.ensure((size_t)count * c->hp.ell * 8);
    c->stage.ensure((size_t)count * poly * 8);
    CUDA_CHECK(cudaMemcpyAsync(c->in_small.p, coeffs, (size_t)count * c->hp.ell * 8, cudaMemcpyHostToDevice, c->stream));
    ntt(c, c->in_small.p, 8, nullptr, count, 1, c->stage.as<u64>(), poly, c->hp.ell);
    CUDA_CHECK(cudaMemcpyAsync(out, c->stage.p, (size_t)count * poly * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
}

int pvw_encode_scalars(pvw_ctx* c, uint32_t count, const uint64_t* m, uint64_t* out) {
  return guarded(c, [&] {
    if (count == 0) return;
    require(m && out, PVW_ERR_INVALID_PARAMETERS, "null argument");
    const size_t poly = c->poly();
    c->in_small.ensure((size_t)count * c->hp.ell * 8);
    c->in_m.ensure((size_t)count * 8);
    c->stage.ensure((size_t)count * poly * 8);
    CUDA_CHECK(cudaMemsetAsync(c->in_small.p, 0, (size_t)count * c->hp.ell * 8, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->in_m.p, m, (size_t)count * 8, cudaMemcpyHostToDevice, c->stream));
    ntt(c, c->in_small.p, 8, c->in_m.as<u64>(), count, 1, c->stage.as<u64>(), poly, c->hp.ell);
    CUDA_CHECK(cudaMemcpyAsync(out, c->stage.p, (size_t)count * poly * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
}

// ------------------------------------------------------------------------------------------------------------
// wire format
// ------------------------------------------------------------------------------------------------------------
namespace {

struct WireSizes { uint64_t rec, row, ct, crs, params; };
WireSizes wire_sizes(const pvw_ctx* c) {
  WireSizes z;
  z.rec = c->W.rec_bytes;
  z.params = c->params_blob.size();
  z.row = 8 + (uint64_t)c->hp.k * z.rec;                                         // Vec<Vec<u8>> of k polynomials
  z.ct = z.row + 8 + (uint64_t)c->hp.n * z.rec + z.params;                       // encryption.rs:298-317
  z.crs = 8 + (uint64_t)c->hp.k * z.row + z.params;                              // crs.rs:228-249
  return z;
}
void wire_pack(pvw_ctx* c, const u64* src, size_t ls, size_t bs, uint64_t count, uint32_t batch, bool packed, uint8_t* out, size_t obs) {
  const double bytes = (double)batch * count * (c->poly() * 8.0 + c->W.rec_bytes);
  launch(c, PVW_KERNEL_WIRE, bytes, [&] { launch_wire_pack(c->W, src, ls, bs, count, batch, packed ? 1 : 0, out, obs, c->stream); });
}
void wire_unpack(pvw_ctx* c, const uint8_t* in, size_t ibs, uint64_t count, uint32_t batch, u64* dst, size_t ls, size_t bs, bool packed, bool write) {
  const double bytes = (double)batch * count * ((write ? c->poly() * 8.0 : 0.0) + c->W.rec_bytes);
  launch(c, PVW_KERNEL_WIRE, bytes, [&] { launch_wire_unpack(c->W, in, ibs, count, batch, dst, ls, bs, packed ? 1 : 0, write ? 1 : 0, c->wire_err, c->stream); });
}
void wire_fill(pvw_ctx* c, uint8_t* out, size_t obs, uint32_t batch, const uint8_t* tmpl, uint32_t len) {
  launch(c, PVW_KERNEL_WIRE, 0.0, [&] { launch_wire_fill(out, obs, batch, tmpl, len, c->stream); });
}
void wire_expect(pvw_ctx* c, const uint8_t* in, size_t ibs, uint32_t batch, const uint8_t* tmpl, uint32_t len) {
  launch(c, PVW_KERNEL_WIRE, 0.0, [&] { launch_wire_expect(in, ibs, batch, tmpl, len, c->wire_err, c->stream); });
}
void wire_err_reset(pvw_ctx* c) { CUDA_CHECK(cudaMemsetAsync(c->wire_err, 0, 4, c->stream)); }
// synchronises; throws DeserializationError when any validation kernel since the last reset flagged something
void wire_err_check(pvw_ctx* c, const char* what) {
  int e = 0;
  CUDA_CHECK(cudaMemcpyAsync(&e, c->wire_err, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (!e) return;
  std::string why;
  if (e & WIRE_ERR_ENVELOPE) why += " length prefixes or embedded parameters differ from this context's;";
  if (e & WIRE_ERR_HEADER) why += " a polynomial record has the wrong length, representation or degree;";
  if (e & WIRE_ERR_RESIDUE) why += " a coefficient is not reduced modulo its prime;";
  throw PvwException(PVW_ERR_DESERIALIZATION, std::string(what) + ":" + why);
}
// bytes -> device: returns a device pointer holding `bytes` bytes of `src` (host or device per flags)
const uint8_t* wire_stage_in(pvw_ctx* c, const uint8_t* src, size_t bytes, uint32_t flags) {
  if (flags & PVW_IO_DEVICE) return src;
  c->wire_buf.ensure(bytes);
  CUDA_CHECK(cudaMemcpyAsync(c->wire_buf.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return c->wire_buf.as<uint8_t>();
}

}  // namespace

int pvw_wire_layout_get(const pvw_ctx* c, pvw_wire_layout* out) {
  if (!c || !out) return PVW_ERR_INVALID_PARAMETERS;
  const WireSizes z = wire_sizes(c);
  out->poly_bytes = c->rq_bytes; out->record_bytes = z.rec; out->params_bytes = z.params; out->pk_row_bytes = z.row;
  out->ciphertext_bytes = z.ct; out->crs_bytes = z.crs;
  out->ct_c1_offset = 0; out->ct_c2_offset = z.row; out->ct_params_offset = z.ct - z.params;
  return PVW_OK;
}
int pvw_wire_params(const pvw_ctx* c, uint8_t* out, uint64_t cap) {
  if (!c || !out || cap < c->params_blob.size()) return PVW_ERR_INVALID_PARAMETERS;
  memcpy(out, c->params_blob.data(), c->params_blob.size());
  return PVW_OK;
}

int pvw_wire_ct_serialize(pvw_ctx* c, uint32_t slot0, uint32_t D, uint8_t* out, uint64_t stride, uint32_t flags) {
  return guarded(c, [&] {
    if (D == 0) return;
    const WireSizes z = wire_sizes(c);
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell, nrows = c->nrows, n = c->hp.n;
    require(out != nullptr, PVW_ERR_INVALID_PARAMETERS, "out is null");
    require(stride >= z.ct, PVW_ERR_INVALID_PARAMETERS, fmt("stride %llu is smaller than one serialised ciphertext (%llu bytes)", (unsigned long long)stride, (unsigned long long)z.ct));
    require((uint64_t)slot0 + D <= c->cap, PVW_ERR_INDEX_OUT_OF_BOUNDS, fmt("ciphertext slots [%u, %u) exceed the reserved capacity %u", slot0, slot0 + D, c->cap));
    const size_t w1 = c->c1_stride(), w2 = (size_t)L * nrows * ell;
    const bool host = !(flags & PVW_IO_DEVICE);
    const uint32_t per = host ? (uint32_t)std::max<uint64_t>(1, (uint64_t)c->upload_chunk_bytes / z.ct) : D;
    for (uint32_t d0 = 0; d0 < D; d0 += per) {
      const uint32_t Dc = std::min(per, D - d0);
      uint8_t* dev = out + (size_t)d0 * stride;
      size_t ds = stride;
      if (host) {
        c->wire_buf.ensure((size_t)Dc * z.ct);
        dev = c->wire_buf.as<uint8_t>(); ds = z.ct;
        if (nrows < n) CUDA_CHECK(cudaMemsetAsync(dev, 0, (size_t)Dc * z.ct, c->stream));  // records of rows held by other shards
      }
      wire_fill(c, dev, ds, Dc, c->env_k, 8);
      wire_pack(c, c->c1s.as<u64>() + (size_t)(slot0 + d0) * w1, (size_t)k * ell, w1, k, Dc, true, dev + 8, ds);
      wire_fill(c, dev + z.row, ds, Dc, c->env_n, 8);
      wire_pack(c, c->c2s.as<u64>() + (size_t)(slot0 + d0) * w2, (size_t)nrows * ell, w2, nrows, Dc, false, dev + z.row + 8 + (size_t)c->row0 * z.rec, ds);
      wire_fill(c, dev + z.ct - z.params, ds, Dc, c->env_params, (uint32_t)z.params);
      if (host) {
        CUDA_CHECK(cudaMemcpy2DAsync(out + (size_t)d0 * stride, stride, dev, z.ct, z.ct, Dc, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
      }
    }
  });
}

int pvw_wire_ct_deserialize(pvw_ctx* c, uint32_t slot0, uint32_t D, const uint8_t* in, uint64_t stride, uint32_t flags) {
  return guarded(c, [&] {
    if (D == 0) return;
    const WireSizes z = wire_sizes(c);
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell, nrows = c->nrows;
    require(in != nullptr, PVW_ERR_INVALID_PARAMETERS, "in is null");
    require(stride >= z.ct, PVW_ERR_INSUFFICIENT_DATA, fmt("expected %llu bytes per ciphertext, got %llu", (unsigned long long)z.ct, (unsigned long long)stride));
    require((uint64_t)slot0 + D <= c->cap, PVW_ERR_INDEX_OUT_OF_BOUNDS, fmt("ciphertext slots [%u, %u) exceed the reserved capacity %u", slot0, slot0 + D, c->cap));
    const size_t w1 = c->c1_stride(), w2 = (size_t)L * nrows * ell;
    c->c1_planes_invalidate(slot0, D);
    const bool host = !(flags & PVW_IO_DEVICE);
    const uint32_t per = host ? (uint32_t)std::max<uint64_t>(1, (uint64_t)c->upload_chunk_bytes / z.ct) : D;
    // Host input larger than one staging chunk: validate EVERY chunk before the first store write, so that a malformed blob
    // anywhere in the batch leaves the store untouched (the reference returns Err without side effects); a single chunk is
    // validated and written in one residency.
    const bool two_pass = D > per;
    for (int pass = two_pass ? 0 : 1; pass < 2; pass++) {
      for (uint32_t d0 = 0; d0 < D; d0 += per) {
        const uint32_t Dc = std::min(per, D - d0);
        const uint8_t* dev = in + (size_t)d0 * stride;
        size_t ds = stride;
        if (host) {
          c->wire_buf.ensure((size_t)Dc * z.ct);
          CUDA_CHECK(cudaMemcpy2DAsync(c->wire_buf.p, z.ct, in + (size_t)d0 * stride, stride, z.ct, Dc, cudaMemcpyHostToDevice, c->stream));
          dev = c->wire_buf.as<uint8_t>(); ds = z.ct;
        }
        u64* c1 = c->c1s.as<u64>() + (size_t)(slot0 + d0) * w1;
        u64* c2 = c->c2s.as<u64>() + (size_t)(slot0 + d0) * w2;
        const uint8_t* r2 = dev + z.row + 8 + (size_t)c->row0 * z.rec;
        if (pass == 0 || !two_pass) {
          wire_err_reset(c);
          wire_expect(c, dev, ds, Dc, c->env_k, 8);
          wire_expect(c, dev + z.row, ds, Dc, c->env_n, 8);
          wire_expect(c, dev + z.ct - z.params, ds, Dc, c->env_params, (uint32_t)z.params);
          wire_unpack(c, dev + 8, ds, k, Dc, c1, (size_t)k * ell, w1, true, false);
          wire_unpack(c, r2, ds, nrows, Dc, c2, (size_t)nrows * ell, w2, false, false);
          wire_err_check(c, "PvwCiphertext");
        }
        if (pass == 1) {
          wire_unpack(c, dev + 8, ds, k, Dc, c1, (size_t)k * ell, w1, true, true);
          wire_unpack(c, r2, ds, nrows, Dc, c2, (size_t)nrows * ell, w2, false, true);
        }
        if (host) CUDA_CHECK(cudaStreamSynchronize(c->stream));  // the staging buffer is reused by the next chunk
      }
    }
  });
}

int pvw_wire_pk_serialize_rows(pvw_ctx* c, uint32_t row, uint32_t count, uint8_t* out, uint32_t flags) {
  return guarded(c, [&] {
    if (count == 0) return;
    require(out != nullptr, PVW_ERR_INVALID_PARAMETERS, "out is null");
    check_rows(c, row, count);
    ensure_B(c);
    const WireSizes z = wire_sizes(c);
    const uint32_t k = c->hp.k, ell = c->hp.ell;
    const bool host = !(flags & PVW_IO_DEVICE);
    const uint32_t per = host ? (uint32_t)std::max<uint64_t>(1, (uint64_t)c->upload_chunk_bytes / z.row) : count;
    for (uint32_t r0 = 0; r0 < count; r0 += per) {
      const uint32_t rc = std::min(per, count - r0);
      uint8_t* dev = out + (size_t)r0 * z.row;
      if (host) { c->wire_buf.ensure((size_t)rc * z.row); dev = c->wire_buf.as<uint8_t>(); }
      wire_fill(c, dev, z.row, rc, c->env_k, 8);
      wire_pack(c, c->B.as<u64>() + (size_t)(row - c->row0 + r0) * k * ell, (size_t)c->nrows * k * ell, (size_t)k * ell, k, rc, true, dev + 8, z.row);
      if (host) {
        CUDA_CHECK(cudaMemcpyAsync(out + (size_t)r0 * z.row, dev, (size_t)rc * z.row, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
      }
    }
  });
}

int pvw_wire_pk_deserialize_rows(pvw_ctx* c, uint32_t row, uint32_t count, const uint8_t* in, uint32_t flags) {
  return guarded(c, [&] {
    if (count == 0) return;
    require(in != nullptr, PVW_ERR_INVALID_PARAMETERS, "in is null");
    check_rows(c, row, count);
    ensure_B(c);
    const WireSizes z = wire_sizes(c);
    const uint32_t k = c->hp.k, ell = c->hp.ell;
    const bool host = !(flags & PVW_IO_DEVICE);
    const uint32_t per = host ? (uint32_t)std::max<uint64_t>(1, (uint64_t)c->upload_chunk_bytes / z.row) : count;
    for (uint32_t r0 = 0; r0 < count; r0 += per) {
      const uint32_t rc = std::min(per, count - r0);
      const uint8_t* dev = wire_stage_in(c, in + (size_t)r0 * z.row, (size_t)rc * z.row, flags);
      u64* dst = c->B.as<u64>() + (size_t)(row - c->row0 + r0) * k * ell;
      wire_err_reset(c);
      wire_expect(c, dev, z.row, rc, c->env_k, 8);
      wire_unpack(c, dev + 8, z.row, k, rc, dst, (size_t)c->nrows * k * ell, (size_t)k * ell, true, false);
      wire_err_check(c, "public key rows");
      wire_unpack(c, dev + 8, z.row, k, rc, dst, (size_t)c->nrows * k * ell, (size_t)k * ell, true, true);
      if (host) CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    c->num_keys = std::max(c->num_keys, row + count); c->Bs_valid = false;
  });
}

int pvw_wire_crs_serialize(pvw_ctx* c, uint8_t* out, uint64_t cap, uint32_t flags) {
  return guarded(c, [&] {
    const WireSizes z = wire_sizes(c);
    require(out != nullptr && cap >= z.crs, PVW_ERR_INVALID_PARAMETERS, fmt("the serialised CRS needs %llu bytes", (unsigned long long)z.crs));
    require(c->A_set, PVW_ERR_INVALID_PARAMETERS, "CRS has not been uploaded");
    const uint32_t k = c->hp.k, ell = c->hp.ell;
    const bool host = !(flags & PVW_IO_DEVICE);
    uint8_t* dev = out;
    if (host) { c->wire_buf.ensure(z.crs); dev = c->wire_buf.as<uint8_t>(); }
    wire_fill(c, dev, 0, 1, c->env_k, 8);                                          // outer Vec: k rows
    wire_fill(c, dev + 8, z.row, k, c->env_k, 8);                                  // each row: k records
    wire_pack(c, c->A.as<u64>(), (size_t)k * k * ell, (size_t)k * ell, k, k, true, dev + 16, z.row);
    wire_fill(c, dev + z.crs - z.params, 0, 1, c->env_params, (uint32_t)z.params);
    if (host) {
      CUDA_CHECK(cudaMemcpyAsync(out, dev, z.crs, cudaMemcpyDeviceToHost, c->stream));
      CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
  });
}

int pvw_wire_crs_deserialize(pvw_ctx* c, const uint8_t* in, uint64_t len, uint32_t flags) {
  return guarded(c, [&] {
    const WireSizes z = wire_sizes(c);
    require(in != nullptr, PVW_ERR_INVALID_PARAMETERS, "in is null");
    require(len >= z.crs, PVW_ERR_INSUFFICIENT_DATA, fmt("expected %llu bytes, got %llu", (unsigned long long)z.crs, (unsigned long long)len));
    const uint32_t L = c->hp.L, k = c->hp.k, ell = c->hp.ell;
    const uint8_t* dev = wire_stage_in(c, in, z.crs, flags);
    c->A.ensure((size_t)L * k * k * ell * 8);
    wire_err_reset(c);
    wire_expect(c, dev, 0, 1, c->env_k, 8);
    wire_expect(c, dev + 8, z.row, k, c->env_k, 8);
    wire_expect(c, dev + z.crs - z.params, 0, 1, c->env_params, (uint32_t)z.params);
    wire_unpack(c, dev + 16, z.row, k, k, c->A.as<u64>(), (size_t)k * k * ell, (size_t)k * ell, true, false);
    wire_err_check(c, "PvwCrs");
    wire_unpack(c, dev + 16, z.row, k, k, c->A.as<u64>(), (size_t)k * k * ell, (size_t)k * ell, true, true);
    c->A_set = true;
    c->At_valid = false; c->As_valid = false;
    if (!(flags & PVW_IO_DEVICE)) CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
}

int pvw_wire_polys_serialize(pvw_ctx* c, uint32_t count, const uint64_t* polys, uint8_t* out) {
  return guarded(c, [&] {
    if (count == 0) return;
    require(polys && out, PVW_ERR_INVALID_PARAMETERS, "null argument");
    const size_t poly = c->poly(), rec = c->W.rec_bytes;
    const uint32_t ell = c->hp.ell;
    c->stage.ensure((size_t)count * poly * 8);
    c->z.ensure((size_t)count * poly * 8);
    c->wire_buf.ensure((size_t)count * rec);
    CUDA_CHECK(cudaMemcpyAsync(c->stage.p, polys, (size_t)count * poly * 8, cudaMemcpyHostToDevice, c->stream));
    to_limb_major(c, c->stage.as<u64>(), count, c->z.as<u64>(), (size_t)count * ell, false);
    wire_pack(c, c->z.as<u64>(), (size_t)count * ell, 0, count, 1, false, c->wire_buf.as<uint8_t>(), 0);
    CUDA_CHECK(cudaMemcpyAsync(out, c->wire_buf.p, (size_t)count * rec, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
}

int pvw_wire_polys_deserialize(pvw_ctx* c, uint32_t count, const uint8_t* in, uint64_t* polys) {
  return guarded(c, [&] {
    if (count == 0) return;
    require(polys && in, PVW_ERR_INVALID_PARAMETERS, "null argument");
    const size_t poly = c->poly(), rec = c->W.rec_bytes;
    const uint32_t ell = c->hp.ell;
    c->stage.ensure((size_t)count * poly * 8);
    c->z.ensure((size_t)count * poly * 8);
    const uint8_t* dev = wire_stage_in(c, in, (size_t)count * rec, PVW_IO_HOST);
    wire_err_reset(c);
    wire_unpack(c, dev, 0, count, 1, c->z.as<u64>(), (size_t)count * ell, 0, false, true);
    wire_err_check(c, "polynomial records");
    from_limb_major(c, c->z.as<u64>(), (size_t)count * ell, count, c->stage.as<u64>(), false);
    CUDA_CHECK(cudaMemcpyAsync(polys, c->stage.p, (size_t)count * poly * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
  });
}

// ------------------------------------------------------------------------------------------------------------
// multi-GPU: exchange of the c1 dealer slices between the row-sharded contexts of one box
// ------------------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*stream_value_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
struct StreamMemOps { stream_value_fn wait = nullptr, write = nullptr; };
const StreamMemOps& stream_mem_ops() {
  static StreamMemOps ops;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      ops.wait = reinterpret_cast<stream_value_fn>(fn);
    if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      ops.write = reinterpret_cast<stream_value_fn>(fn);
  }
  return ops;
}
void stream_wait_geq(cudaStream_t st, const u64* addr, u64 value) {
  const CUresult r = stream_mem_ops().wait((CUstream)st, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) throw PvwException(PVW_ERR_INTERNAL, fmt("cuStreamWaitValue64 failed (CUresult %d)", (int)r));
}
void stream_write(cudaStream_t st, u64* addr, u64 value) {
  const CUresult r = stream_mem_ops().write((CUstream)st, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WRITE_VALUE_DEFAULT);
  if (r != CUDA_SUCCESS) throw PvwException(PVW_ERR_INTERNAL, fmt("cuStreamWriteValue64 failed (CUresult %d)", (int)r));
}

struct ShardBlob {   // what pvw_shard_export hands to the peers (fits pvw_shard_handle)
  cudaIpcMemHandle_t c1, flags;
  uint64_t cap, w1, world;
  uint64_t magic;
};
static_assert(sizeof(ShardBlob) <= sizeof(pvw_shard_handle), "pvw_shard_handle too small");
constexpr uint64_t SHARD_MAGIC = 0x5056573153484452ull;

}  // namespace

static void shard_disconnect(pvw_ctx* c) {
  pvw_ctx::Shard& sh = c->sh;
  if (!sh.connected) return;
  if (sh.xstream) cudaStreamSynchronize(sh.xstream);
  cudaStreamSynchronize(c->stream);
  for (uint32_t r = 0; r < sh.world; r++) {
    if (r == sh.rank) continue;
    if (sh.peer_c1[r]) cudaIpcCloseMemHandle(sh.peer_c1[r]);
    if (sh.peer_flags[r]) cudaIpcCloseMemHandle(sh.peer_flags[r]);
  }
  sh.peer_c1.clear(); sh.peer_flags.clear();
  sh.connected = false;
}

int pvw_shard_export(pvw_ctx* c, uint32_t world, pvw_shard_handle* out) {
  return guarded(c, [&] {
    require(out != nullptr && world >= 1, PVW_ERR_INVALID_PARAMETERS, "null argument / empty world");
    require(c->cap > 0, PVW_ERR_INVALID_PARAMETERS, "pvw_shard_export: reserve the ciphertext store first (pvw_ct_reserve)");
    require(!c->sh.connected, PVW_ERR_INVALID_PARAMETERS, "pvw_shard_export: already connected; disconnect first");
    require(stream_mem_ops().wait && stream_mem_ops().write, PVW_ERR_INTERNAL, "this driver does not export cuStreamWaitValue64 / cuStreamWriteValue64");
    pvw_ctx::Shard& sh = c->sh;
    if (!sh.xstream) CUDA_CHECK(cudaStreamCreateWithFlags(&sh.xstream, cudaStreamNonBlocking));
    if (!sh.ev_c1) CUDA_CHECK(cudaEventCreateWithFlags(&sh.ev_c1, cudaEventDisableTiming));
    if (sh.flags && sh.flags_world != world) { CUDA_CHECK(cudaFree(sh.flags)); sh.flags = nullptr; }
    if (!sh.flags) { CUDA_CHECK(cudaMalloc(&sh.flags, (2 * (size_t)world + 2) * 8)); sh.flags_world = world; }
    CUDA_CHECK(cudaMemset(sh.flags, 0, (2 * (size_t)world + 2) * 8));   // counters restart with every (re)connection
    sh.push_seq = sh.wait_seq = 0;
    ShardBlob b;
    memset(&b, 0, sizeof(b));
    CUDA_CHECK(cudaIpcGetMemHandle(&b.c1, c->c1s.p));
    CUDA_CHECK(cudaIpcGetMemHandle(&b.flags, sh.flags));
    b.cap = c->cap; b.w1 = c->c1_stride(); b.world = world; b.magic = SHARD_MAGIC;
    memset(out, 0, sizeof(*out));
    memcpy(out, &b, sizeof(b));
  });
}

int pvw_shard_connect(pvw_ctx* c, uint32_t world, uint32_t rank, const pvw_shard_handle* all) {
  return guarded(c, [&] {
    require(all != nullptr && world >= 1 && rank < world, PVW_ERR_INVALID_PARAMETERS, "bad world / rank");
    pvw_ctx::Shard& sh = c->sh;
    require(!sh.connected, PVW_ERR_INVALID_PARAMETERS, "pvw_shard_connect: already connected");
    require(sh.flags && sh.flags_world == world, PVW_ERR_INVALID_PARAMETERS, "pvw_shard_connect: call pvw_shard_export(world) first");
    const uint64_t w1 = c->c1_stride();
    sh.peer_c1.assign(world, nullptr); sh.peer_flags.assign(world, nullptr);
    sh.world = world; sh.rank = rank;
    sh.connected = true;                       // from here on a failure is cleaned up by shard_disconnect
    try {
      for (uint32_t r = 0; r < world; r++) {
        ShardBlob b;
        memcpy(&b, &all[r], sizeof(b));
        require(b.magic == SHARD_MAGIC && b.world == world, PVW_ERR_INVALID_PARAMETERS, fmt("handle of rank %u is not a pvw_shard_handle of this world", r));
        require(b.cap == c->cap && b.w1 == w1, PVW_ERR_DIMENSION_MISMATCH,
                fmt("rank %u reserved %llu ciphertext slots of %llu words, this rank %u of %llu", r, (unsigned long long)b.cap, (unsigned long long)b.w1, c->cap, (unsigned long long)w1));
        if (r == rank) { sh.peer_c1[r] = c->c1s.as<u64>(); sh.peer_flags[r] = sh.flags; continue; }
        void* p = nullptr;
        CUDA_CHECK(cudaIpcOpenMemHandle(&p, b.c1, cudaIpcMemLazyEnablePeerAccess));
        sh.peer_c1[r] = reinterpret_cast<u64*>(p);
        CUDA_CHECK(cudaIpcOpenMemHandle(&p, b.flags, cudaIpcMemLazyEnablePeerAccess));
        sh.peer_flags[r] = reinterpret_cast<u64*>(p);
      }
    } catch (...) {
      shard_disconnect(c);
      throw;
    }
  });
}

int pvw_shard_disconnect(pvw_ctx* c) {
  return guarded(c, [&] { shard_disconnect(c); });
}

int pvw_shard_push_c1(pvw_ctx* c, uint32_t slot0, uint32_t count) {
  return guarded(c, [&] { shard_push_c1(c, slot0, count); });
}

static void shard_push_c1(pvw_ctx* c, uint32_t slot0, uint32_t count) {
  {
    pvw_ctx::Shard& sh = c->sh;
    require(sh.connected, PVW_ERR_INVALID_PARAMETERS, "pvw_shard_push_c1: not connected");
    require((uint64_t)slot0 + count <= c->cap, PVW_ERR_INDEX_OUT_OF_BOUNDS, fmt("ciphertext slots [%u, %u) exceed the reserved capacity %u", slot0, slot0 + count, c->cap));
    const size_t w1 = c->c1_stride();   // residues and byte planes of a slot travel together
    sh.batch_D = 0;                     // (set again by the encrypt call when the push came from PVW_ENC_PUSH_C1)
    const u64 seq = ++sh.push_seq;
    u64* stage = sh.flags + 2 * (size_t)sh.world;
    // the slice is final once everything queued on the compute stream so far (its c1 product) has run
    CUDA_CHECK(cudaEventRecord(sh.ev_c1, c->stream));
    CUDA_CHECK(cudaStreamWaitEvent(sh.xstream, sh.ev_c1, 0));
    for (uint32_t i = 1; i < sh.world; i++) {
      const uint32_t r = (sh.rank + i) % sh.world;                       // staggered: at any moment every peer receives from one rank
      stream_wait_geq(sh.xstream, sh.flags + sh.world + r, seq - 1);     // peer r has released the c1 of the previous exchange
      if (count)
        CUDA_CHECK(cudaMemcpyAsync(sh.peer_c1[r] + (size_t)slot0 * w1, c->c1s.as<u64>() + (size_t)slot0 * w1, (size_t)count * w1 * 8, cudaMemcpyDeviceToDevice, sh.xstream));
    }
    stream_write(sh.xstream, stage, seq);
    for (uint32_t i = 1; i < sh.world; i++) {
      const uint32_t r = (sh.rank + i) % sh.world;
      CUDA_CHECK(cudaMemcpyAsync(sh.peer_flags[r] + sh.rank, stage, 8, cudaMemcpyDeviceToDevice, sh.xstream));   // "my slice of exchange seq has landed"
    }
  }
}

int pvw_shard_wait_c1(pvw_ctx* c) {
  return guarded(c, [&] {
    pvw_ctx::Shard& sh = c->sh;
    require(sh.connected, PVW_ERR_INVALID_PARAMETERS, "pvw_shard_wait_c1: not connected");
    const u64 seq = ++sh.wait_seq;
    require(seq <= sh.push_seq, PVW_ERR_INVALID_PARAMETERS, "pvw_shard_wait_c1 without a matching pvw_shard_push_c1 on this rank");
    for (uint32_t r = 0; r < sh.world; r++)
      if (r != sh.rank) stream_wait_geq(c->stream, sh.flags + r, seq);
    // the peers' slices of the batch have landed (in stream order): their byte planes are as current as this rank's own
    for (uint32_t i = sh.batch_slot0; i < sh.batch_slot0 + sh.batch_D && i < c->c1p_valid.size(); i++) c->c1p_valid[i] = sh.batch_direct ? 1 : 0;
  });
}

int pvw_shard_release_c1(pvw_ctx* c) {
  return guarded(c, [&] {
    pvw_ctx::Shard& sh = c->sh;
    require(sh.connected, PVW_ERR_INVALID_PARAMETERS, "pvw_shard_release_c1: not connected");
    u64* stage = sh.flags + 2 * (size_t)sh.world + 1;
    stream_write(c->stream, stage, sh.wait_seq);
    for (uint32_t i = 1; i < sh.world; i++) {
      const uint32_t r = (sh.rank + i) % sh.world;
      CUDA_CHECK(cudaMemcpyAsync(sh.peer_flags[r] + sh.world + sh.rank, stage, 8, cudaMemcpyDeviceToDevice, c->stream));
    }
  });
}

int pvw_ctx_synchronize(pvw_ctx* c) {
  return guarded(c, [&] {
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->sh.xstream) CUDA_CHECK(cudaStreamSynchronize(c->sh.xstream));
  });
}
void* pvw_ctx_stream(pvw_ctx* c) { return c ? (void*)c->stream : nullptr; }

int pvw_ctx_set_option(pvw_ctx* c, const char* name, int64_t value) {
  return guarded(c, [&] {
    require(name != nullptr, PVW_ERR_INVALID_PARAMETERS, "null option name");
    std::string n(name);
    if (n == "imma") c->use_imma = value != 0;
    else if (n == "imma_pair") c->imma_pair = value != 0;
    else if (n == "narrow_inputs") c->narrow_inputs = value != 0;
    else if (n == "ternary_tables") { c->use_tern = value != 0; c->T.tern = c->use_tern ? c->tern_dev : nullptr; }
    else if (n == "imma_fast_reduce") c->imma_fast_reduce = value != 0;
    else if (n == "imma_epilogue_warps") { require(value == 0 || value == 8 || value == 16, PVW_ERR_INVALID_PARAMETERS, "imma_epilogue_warps must be 0 (default: 8), 8 or 16"); c->imma_epi_warps = (int)value; }
    else if (n == "planes_only") c->planes_only = value != 0;
    else if (n == "imma_stages") { require(value == 0 || (value >= 2 && value <= 10), PVW_ERR_INVALID_PARAMETERS, "imma_stages must be 0 (auto) or 2..10"); c->imma_stages = (int)value; }
    else if (n == "imma_min_dealers") { require(value >= 1, PVW_ERR_INVALID_PARAMETERS, "imma_min_dealers must be >= 1"); c->imma_min_dealers = value; }
    else if (n == "imma_min_rows") { require(value >= 1, PVW_ERR_INVALID_PARAMETERS, "imma_min_rows must be >= 1"); c->imma_min_rows = value; }
    else if (n == "imma_chunk_dealers") { require(value >= 16 && value <= (1 << 20), PVW_ERR_INVALID_PARAMETERS, "imma_chunk_dealers must be in [16, 2^20]"); c->imma_chunk_dealers = value; }
    else if (n == "gemm_impl") { require(value >= 0 && value <= 2, PVW_ERR_INVALID_PARAMETERS, "gemm_impl must be 0, 1 or 2"); c->gemm_impl = (int)value; }
    else if (n == "gemm_tile") { require(value >= 0 && value <= 3, PVW_ERR_INVALID_PARAMETERS, "gemm_tile must be 0..3"); c->gemm_tile = (int)value; }
    else if (n == "refill_lag") { require(value >= 1 && value <= 5, PVW_ERR_INVALID_PARAMETERS, "refill_lag must be 1..5"); c->refill_lag = (int)value; }
    else if (n == "lift_fast") { require(value == 0 || value == 1, PVW_ERR_INVALID_PARAMETERS, "lift_fast must be 0 or 1"); c->T.lift_fast = (value && c->hp.shortL > 0) ? 1 : 0; }
    else if (n == "decode_fused") { require(value == 0 || value == 1, PVW_ERR_INVALID_PARAMETERS, "decode_fused must be 0 (chain for every share) or 1 (fused fast path)"); c->decode_fused = (int)value; }
    else if (n == "tail_impl") { require(value == 0 || value == 1, PVW_ERR_INVALID_PARAMETERS, "tail_impl must be 0 or 1"); c->T.tail_impl = (int)value; }
    else if (n == "decrypt_chunk_shares") { require(value > 0, PVW_ERR_INVALID_PARAMETERS, "decrypt_chunk_shares must be positive"); c->decrypt_chunk_shares = value; }
    else if (n == "upload_chunk_bytes") { require(value >= 4096, PVW_ERR_INVALID_PARAMETERS, "upload_chunk_bytes too small"); c->upload_chunk_bytes = value; }
    else if (n == "profile") {
      prof_drain(c);
      c->profile = value != 0;
      if (value == 2 || value == 0) for (int i = 0; i < PVW_KERNEL_KINDS; i++) { c->prof_ms[i] = 0; c->prof_n[i] = 0; c->prof_bytes[i] = 0; }
    }
    else throw PvwException(PVW_ERR_INVALID_PARAMETERS, "unknown option " + n);
  });
}
int pvw_ctx_profile(pvw_ctx* c, int kind, double* ms_total, uint64_t* launches, double* algorithmic_bytes) {
  return guarded(c, [&] {
    require(kind >= 0 && kind < PVW_KERNEL_KINDS, PVW_ERR_INVALID_PARAMETERS, "unknown kernel kind");
    prof_drain(c);
    if (ms_total) *ms_total = c->prof_ms[kind];
    if (launches) *launches = c->prof_n[kind];
    if (algorithmic_bytes) *algorithmic_bytes = c->prof_bytes[kind];
  });
}
uint64_t pvw_ctx_launch_count(const pvw_ctx* c) { return c ? c->launches : 0; }
const char* pvw_version(void) { return "pvw_b200 0.2 (sm_100a; tcgen05 int8 product)"; }

}  // extern "C"
