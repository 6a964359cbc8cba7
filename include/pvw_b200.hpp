// pvw_b200.hpp -- C++ host-side mirror of the pvw-rs API for the hot path, header-only, over the C ABI of pvw_b200.h.
//
// The reference is a compiled (Rust) crate whose toolchain is absent from the build image, so this is the compiled-code
// host layer: same names, argument meaning and error behaviour as the crate's public API (src/lib.rs:31-55), every ring
// operation forwarded to libpvw_b200.so.  Polynomials are flat u64 blocks in the reference layout ([L][ell] row-major,
// NTT form); matrices / vectors of polynomials are std::vector<uint64_t>.  The Rust shim of INTEGRATION.md is the same
// code with `Poly` at the edges.
#pragma once
#include <sys/random.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "pvw_b200.h"

namespace pvw {

// PvwError (src/errors.rs:11-73): variant() names the reference variant the status code maps to
class PvwError : public std::runtime_error {
 public:
  PvwError(int status, const std::string& msg) : std::runtime_error(variant_name(status) + ": " + msg), status_(status) {}
  int status() const { return status_; }
  std::string variant() const { return variant_name(status_); }
  static std::string variant_name(int s) {
    switch (s) {
      case PVW_ERR_INVALID_PARAMETERS: return "InvalidParameters";
      case PVW_ERR_DIMENSION_MISMATCH: return "DimensionMismatch";
      case PVW_ERR_INDEX_OUT_OF_BOUNDS: return "IndexOutOfBounds";
      case PVW_ERR_ENCRYPTION: return "EncryptionError";
      case PVW_ERR_DECRYPTION: return "DecryptionError";
      case PVW_ERR_KEYGEN: return "KeyGenerationError";
      case PVW_ERR_DESERIALIZATION: return "DeserializationError";
      case PVW_ERR_INSUFFICIENT_DATA: return "InsufficientData";
      default: return "InternalError";
    }
  }
 private:
  int status_;
};

// The reference samples secret keys, encryption randomness and errors from `thread_rng()`: `RngCore + CryptoRng`, ChaCha12 keyed
// from the operating system (secret_key.rs:45-63, encryption.rs:138,164,180).  CryptoRng is the equivalent here: ChaCha20 keyed with
// 32 bytes of getrandom(2), usable wherever a C++ UniformRandomBitGenerator is.  A general-purpose PRNG (std::mt19937_64, ...)
// must never produce key material -- its state is recoverable from outputs -- so the samplers below accept nothing else.  Use one
// generator for public values (PvwCrs::new_random) and another for secrets.
class CryptoRng {
 public:
  using result_type = uint64_t;
  static constexpr result_type min() { return 0; }
  static constexpr result_type max() { return ~0ull; }
  CryptoRng() {
    uint8_t key[32];
    size_t got = 0;
    while (got < sizeof(key)) {
      const ssize_t r = getrandom(key + got, sizeof(key) - got, 0);
      if (r < 0) throw std::runtime_error("getrandom failed: no entropy source for key material");
      got += (size_t)r;
    }
    init(key);
    std::memset(key, 0, sizeof(key));
  }
  // Deterministic stream for reproducible tests and fixtures ONLY (the reference offers no such constructor on its samplers).
  static CryptoRng from_seed_for_tests(const uint8_t (&key)[32]) { CryptoRng g(0); g.init(key); return g; }
  CryptoRng(const CryptoRng&) = delete;             // two copies would emit the same "random" secrets
  CryptoRng& operator=(const CryptoRng&) = delete;
  CryptoRng(CryptoRng&&) = default;
  result_type operator()() {
    if (pos_ == 8) refill();
    return buf_[pos_++];
  }
 private:
  explicit CryptoRng(int) {}
  void init(const uint8_t (&key)[32]) {
    static const uint32_t sigma[4] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};   // "expand 32-byte k"
    for (int i = 0; i < 4; i++) st_[i] = sigma[i];
    for (int i = 0; i < 8; i++) st_[4 + i] = (uint32_t)key[4 * i] | (uint32_t)key[4 * i + 1] << 8 | (uint32_t)key[4 * i + 2] << 16 | (uint32_t)key[4 * i + 3] << 24;
    st_[12] = st_[13] = st_[14] = st_[15] = 0;       // 64-bit block counter, 64-bit nonce 0 (one stream per key)
    pos_ = 8;
  }
  static uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
  static void qr(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    a += b; d ^= a; d = rotl(d, 16); c += d; b ^= c; b = rotl(b, 12); a += b; d ^= a; d = rotl(d, 8); c += d; b ^= c; b = rotl(b, 7);
  }
  void refill() {                                    // one ChaCha20 block = eight 64-bit outputs
    uint32_t x[16];
    for (int i = 0; i < 16; i++) x[i] = st_[i];
    for (int r = 0; r < 10; r++) {
      qr(x[0], x[4], x[8], x[12]); qr(x[1], x[5], x[9], x[13]); qr(x[2], x[6], x[10], x[14]); qr(x[3], x[7], x[11], x[15]);
      qr(x[0], x[5], x[10], x[15]); qr(x[1], x[6], x[11], x[12]); qr(x[2], x[7], x[8], x[13]); qr(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 8; i++) buf_[i] = (uint64_t)(x[2 * i] + st_[2 * i]) | (uint64_t)(x[2 * i + 1] + st_[2 * i + 1]) << 32;
    if (++st_[12] == 0) ++st_[13];
    pos_ = 0;
  }
  uint32_t st_[16];
  uint64_t buf_[8];
  int pos_ = 8;
};
using Rng = CryptoRng;

// sample_vec_cbd (src/sampling/uniform.rs:27-70) and sample_uniform_coefficients (:5-22), host side
inline std::vector<int64_t> sample_vec_cbd(size_t n, float variance, Rng& rng) {
  if (!(variance > 0)) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "The variance should be positive");
  std::vector<int64_t> out(n);
  if (std::fabs(variance - 0.5f) < 1.2e-7f) {
    for (auto& v : out) { uint64_t u = rng(); v = (int64_t)(u & 1) - (int64_t)((u >> 1) & 1); }
    return out;
  }
  int v = (int)variance;
  if (v < 1 || v > 16 || (float)v != variance) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "The variance should be an integer between 1 and 16");
  const uint64_t mask = (1ull << (2 * v)) - 1;
  for (auto& x : out) { uint64_t u = rng(); x = (int64_t)__builtin_popcountll(u & mask) - (int64_t)__builtin_popcountll((u >> (2 * v)) & mask); }
  return out;
}
inline std::vector<int64_t> sample_uniform_coefficients(uint64_t bound, size_t n, Rng& rng) {
  std::uniform_int_distribution<int64_t> d(-(int64_t)bound, (int64_t)bound);
  std::vector<int64_t> out(n);
  for (auto& v : out) v = d(rng);
  return out;
}

class Context;  // one pvw_ctx

// PvwParameters (src/params/parameters.rs:19-40) + builder (:44-201)
struct PvwParameters {
  uint32_t n = 0, t = 0, k = 0, l = 0;
  float secret_variance = 0.5f;
  uint64_t error_bound_1 = 100, error_bound_2 = 200;   // parameters.rs:166-168
  std::vector<uint64_t> moduli_, psi;
  std::vector<uint64_t> delta, delta_power_l_minus_1, q_total;  // little-endian words
  int device = 0;
  size_t L() const { return moduli_.size(); }
  size_t poly_words() const { return L() * l; }
  const std::vector<uint64_t>& moduli() const { return moduli_; }
  bool verify_correctness_condition() const;   // parameters.rs:510-551 (evaluated by the library)
  std::shared_ptr<Context> probe;              // validation / derived constants live in a context
};

class Context {
 public:
  Context(const PvwParameters& p, uint32_t row0 = 0, uint32_t nrows = 0) {
    pvw_params_desc d{};
    d.n = p.n; d.k = p.k; d.ell = p.l; d.L = (uint32_t)p.moduli_.size();
    d.moduli = p.moduli_.data(); d.psi = p.psi.empty() ? nullptr : p.psi.data();
    d.secret_variance = p.secret_variance; d.error_bound_1 = p.error_bound_1; d.error_bound_2 = p.error_bound_2;
    d.row0 = row0; d.nrows = nrows; d.device = p.device;
    int rc = pvw_ctx_create(&ctx_, &d);
    if (rc != PVW_OK) throw PvwError(rc, pvw_last_error(nullptr));
  }
  ~Context() { pvw_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  pvw_ctx* get() const { return ctx_; }
  void check(int rc) const { if (rc != PVW_OK) throw PvwError(rc, pvw_last_error(ctx_)); }
  std::vector<uint64_t> bigint(int which) const {
    uint32_t nw = 0;
    check(pvw_params_bigint(ctx_, which, nullptr, 0, &nw));
    std::vector<uint64_t> w(nw ? nw : 1, 0);
    check(pvw_params_bigint(ctx_, which, w.data(), (uint32_t)w.size(), &nw));
    w.resize(nw);
    return w;
  }
  std::mutex mu;   // a context is single-threaded; the reference types are Send + Sync
 private:
  pvw_ctx* ctx_ = nullptr;
};

inline bool PvwParameters::verify_correctness_condition() const {
  int ok = 0;
  probe->check(pvw_params_correctness_condition(probe->get(), &ok));
  return ok != 0;
}

class PvwParametersBuilder {
 public:
  PvwParametersBuilder& set_parties(uint32_t n) { p_.n = n; return *this; }
  PvwParametersBuilder& set_dimension(uint32_t k) { p_.k = k; return *this; }
  PvwParametersBuilder& set_l(uint32_t l) { p_.l = l; return *this; }
  PvwParametersBuilder& set_moduli(const std::vector<uint64_t>& m) { p_.moduli_ = m; return *this; }
  PvwParametersBuilder& set_secret_variance(float v) { p_.secret_variance = v; return *this; }
  PvwParametersBuilder& set_error_bounds_u32(uint32_t b1, uint32_t b2) { p_.error_bound_1 = b1; p_.error_bound_2 = b2; return *this; }
  PvwParametersBuilder& set_psi(const std::vector<uint64_t>& psi) { p_.psi = psi; return *this; }
  PvwParametersBuilder& set_device(int d) { p_.device = d; return *this; }
  // build (parameters.rs:117-195): validation and Delta = floor(Q^(1/l)) are done by pvw_ctx_create
  std::shared_ptr<PvwParameters> build_arc() {
    auto p = std::make_shared<PvwParameters>(p_);
    if (p->moduli_.empty()) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "moduli not set");
    p->probe = std::make_shared<Context>(*p, 0, p->n ? 1 : 0);
    p->t = (p->n - 1) / 2;                                        // parameters.rs:169
    p->q_total = p->probe->bigint(0); p->delta = p->probe->bigint(1); p->delta_power_l_minus_1 = p->probe->bigint(2);
    p->psi.assign(p->moduli_.size(), 0);
    p->probe->check(pvw_params_psi(p->probe->get(), p->psi.data()));
    return p;
  }
 private:
  PvwParameters p_;
};

// PvwCrs (src/params/crs.rs:12-17): k x k polynomials in NTT form, host copy [k][k][L][l]
struct PvwCrs {
  std::shared_ptr<PvwParameters> params;
  std::vector<uint64_t> matrix;
  // crs.rs:24-39: uniform residues labelled NTT
  static PvwCrs new_random(const std::shared_ptr<PvwParameters>& p, Rng& rng) {
    PvwCrs c{p, std::vector<uint64_t>((size_t)p->k * p->k * p->poly_words())};
    for (size_t e = 0; e < (size_t)p->k * p->k; e++)
      for (size_t j = 0; j < p->L(); j++) {
        std::uniform_int_distribution<uint64_t> d(0, p->moduli_[j] - 1);
        for (uint32_t c2 = 0; c2 < p->l; c2++) c.matrix[(e * p->L() + j) * p->l + c2] = d(rng);
      }
    return c;
  }
};

// SecretKey (src/keys/secret_key.rs:14-18): k x l small coefficients
struct SecretKey {
  std::shared_ptr<PvwParameters> params;
  std::vector<int64_t> secret_coeffs;   // [k][l]
  static SecretKey random(const std::shared_ptr<PvwParameters>& p, Rng& rng) {   // secret_key.rs:45-63
    return SecretKey{p, sample_vec_cbd((size_t)p->k * p->l, p->secret_variance, rng)};
  }
};
struct Party {   // public_key.rs:17-27,62-79
  uint32_t index;
  SecretKey sk;
  static Party make(uint32_t index, const std::shared_ptr<PvwParameters>& p, Rng& rng) {
    if (index >= p->n) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "Party index " + std::to_string(index) + " exceeds maximum " + std::to_string(p->n - 1));
    return Party{index, SecretKey::random(p, rng)};
  }
  const SecretKey& secret_key() const { return sk; }
};

// GlobalPublicKey (src/keys/public_key.rs:43-54): B (n x k, NTT) and the CRS A are device resident in one context
class GlobalPublicKey {
 public:
  std::shared_ptr<PvwParameters> params;
  explicit GlobalPublicKey(const PvwCrs& crs) : params(crs.params), ctx_(std::make_shared<Context>(*crs.params)) {
    ctx_->check(pvw_crs_upload(ctx_->get(), crs.matrix.data(), PVW_IO_HOST));
    ctx_->check(pvw_ct_reserve(ctx_->get(), params->n));
  }
  uint32_t num_public_keys() const { uint32_t v = 0; ctx_->check(pvw_pk_num_keys(ctx_->get(), &v)); return v; }
  bool is_full() const { return num_public_keys() >= params->n; }                       // public_key.rs:349-351
  // generate_and_add_party (public_key.rs:256-263): b = s*A + e generated on the device straight into row `index`
  void generate_and_add_party(const Party& party, Rng& rng) {
    auto e = sample_uniform_coefficients(params->error_bound_1, (size_t)params->k * params->l, rng);
    std::lock_guard<std::mutex> g(ctx_->mu);
    ctx_->check(pvw_keygen_batch(ctx_->get(), party.index, 1, party.sk.secret_coeffs.data(), e.data(), PVW_IO_HOST));
  }
  // generate_all_party_keys (public_key.rs:376-401): one batched device call
  // every key lands in row party.index (add_public_key, public_key.rs:214-250) whatever the order of the list; one device call
  // per run of consecutive indices (a single call for the usual 0..len-1 list)
  void generate_all_party_keys(const std::vector<Party>& parties, Rng& rng) {
    if (parties.size() > params->n) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "Too many parties: " + std::to_string(parties.size()) + " > " + std::to_string(params->n));
    const size_t w = (size_t)params->k * params->l;
    std::vector<int64_t> sk(parties.size() * w);
    for (size_t i = 0; i < parties.size(); i++) {
      if (parties[i].index >= params->n) throw PvwError(PVW_ERR_INDEX_OUT_OF_BOUNDS, "Party index " + std::to_string(parties[i].index) + " exceeds maximum " + std::to_string(params->n - 1));
      std::copy(parties[i].sk.secret_coeffs.begin(), parties[i].sk.secret_coeffs.end(), sk.begin() + i * w);
    }
    auto e = sample_uniform_coefficients(params->error_bound_1, parties.size() * w, rng);
    std::lock_guard<std::mutex> g(ctx_->mu);
    for (size_t start = 0; start < parties.size();) {
      size_t stop = start + 1;
      while (stop < parties.size() && parties[stop].index == parties[stop - 1].index + 1) stop++;
      ctx_->check(pvw_keygen_batch(ctx_->get(), parties[start].index, (uint32_t)(stop - start), sk.data() + start * w, e.data() + start * w, PVW_IO_HOST));
      start = stop;
    }
  }
  std::vector<uint64_t> get_public_key(uint32_t index) const {                         // k polynomials of row `index`
    std::vector<uint64_t> row((size_t)params->k * params->poly_words());
    std::lock_guard<std::mutex> g(ctx_->mu);
    ctx_->check(pvw_pk_download_rows(ctx_->get(), index, 1, row.data()));
    return row;
  }
  const std::shared_ptr<Context>& context() const { return ctx_; }
 private:
  std::shared_ptr<Context> ctx_;
};

// PvwCiphertext (src/crypto/encryption.rs:15-24), resident in store slot `slot` of the key's context
struct PvwCiphertext {
  std::shared_ptr<PvwParameters> params;
  std::shared_ptr<Context> ctx;
  uint32_t slot;
  std::vector<uint64_t> c1() const { std::vector<uint64_t> v((size_t)params->k * params->poly_words()); std::lock_guard<std::mutex> g(ctx->mu); ctx->check(pvw_ct_download(ctx->get(), slot, v.data(), nullptr)); return v; }
  std::vector<uint64_t> c2() const { std::vector<uint64_t> v((size_t)params->n * params->poly_words()); std::lock_guard<std::mutex> g(ctx->mu); ctx->check(pvw_ct_download(ctx->get(), slot, nullptr, v.data())); return v; }
  size_t len() const { return params->n; }
};

namespace detail {
inline std::vector<PvwCiphertext> encrypt_many(const std::vector<uint64_t>& m /* [D][n] */, uint32_t D, const GlobalPublicKey& pk, Rng& rng) {
  const auto& p = *pk.params;
  // r (CBD), e1, e2 (uniform) sampled exactly where the reference samples them (encryption.rs:135-142,161-167,196)
  auto r = sample_vec_cbd((size_t)D * p.k * p.l, p.secret_variance, rng);
  auto e1 = sample_uniform_coefficients(p.error_bound_1, (size_t)D * p.k * p.l, rng);
  auto e2 = sample_uniform_coefficients(p.error_bound_2, (size_t)D * p.n * p.l, rng);
  auto ctx = pk.context();
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->check(pvw_encrypt_batch(ctx->get(), 0, D, 0, D, m.data(), r.data(), e1.data(), e2.data(), PVW_IO_HOST));
  std::vector<PvwCiphertext> out;
  for (uint32_t d = 0; d < D; d++) out.push_back(PvwCiphertext{pk.params, ctx, d});
  return out;
}
}  // namespace detail

// encrypt (src/crypto/encryption.rs:105-214); the ciphertext occupies store slot 0 of the key's context
inline PvwCiphertext encrypt(const std::vector<uint64_t>& scalars, const GlobalPublicKey& pk, Rng& rng) {
  if (scalars.size() != pk.params->n)
    throw PvwError(PVW_ERR_INVALID_PARAMETERS, "Must provide exactly n=" + std::to_string(pk.params->n) + " scalars, got " + std::to_string(scalars.size()));
  return detail::encrypt_many(scalars, 1, pk, rng)[0];
}
// encrypt_party_shares (encryption.rs:221-245)
inline PvwCiphertext encrypt_party_shares(const std::vector<uint64_t>& shares, size_t party_index, const GlobalPublicKey& pk, Rng& rng) {
  if (party_index >= pk.params->n) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "Party index " + std::to_string(party_index) + " exceeds maximum " + std::to_string(pk.params->n - 1));
  if (shares.size() != pk.params->n) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "Party must provide " + std::to_string(pk.params->n) + " shares, got " + std::to_string(shares.size()));
  return encrypt(shares, pk, rng);
}
// encrypt_all_party_shares (encryption.rs:253-286): all n dealers in ONE device call (the reference fans out with rayon)
inline std::vector<PvwCiphertext> encrypt_all_party_shares(const std::vector<std::vector<uint64_t>>& all_shares, const GlobalPublicKey& pk, Rng& rng) {
  const uint32_t n = pk.params->n;
  if (all_shares.size() != n) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "Must provide shares for all " + std::to_string(n) + " parties");
  std::vector<uint64_t> m((size_t)n * n);
  for (uint32_t d = 0; d < n; d++) {
    if (all_shares[d].size() != n)
      throw PvwError(PVW_ERR_INVALID_PARAMETERS, "Dealer " + std::to_string(d) + " provided " + std::to_string(all_shares[d].size()) + " shares but needs " + std::to_string(n));
    std::copy(all_shares[d].begin(), all_shares[d].end(), m.begin() + (size_t)d * n);
  }
  return detail::encrypt_many(m, n, pk, rng);
}
// encrypt_broadcast (encryption.rs:292-296)
inline PvwCiphertext encrypt_broadcast(uint64_t scalar, const GlobalPublicKey& pk, Rng& rng) {
  return encrypt(std::vector<uint64_t>(pk.params->n, scalar), pk, rng);
}

// decrypt_party_value (src/crypto/decryption.rs:249-278)
inline uint64_t decrypt_party_value(const PvwCiphertext& ct, const SecretKey& sk, size_t party_index) {
  if (party_index >= ct.params->n) throw PvwError(PVW_ERR_INDEX_OUT_OF_BOUNDS, "party index out of range");  // the reference panics here (:274)
  uint32_t slot = ct.slot, pidx = (uint32_t)party_index;
  uint64_t out = 0;
  std::lock_guard<std::mutex> g(ct.ctx->mu);
  ct.ctx->check(pvw_decrypt_batch(ct.ctx->get(), 1, &slot, 1, &pidx, sk.secret_coeffs.data(), &out, PVW_IO_HOST));
  return out;
}
// decrypt_party_shares (decryption.rs:281-325): exactly n ciphertexts, one batched device call over the dealers
inline std::vector<uint64_t> decrypt_party_shares(const std::vector<PvwCiphertext>& cts, const SecretKey& sk, size_t party_index) {
  if (cts.empty()) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "No ciphertexts provided");
  const auto& p = *cts[0].params;
  if (cts.size() != p.n) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "Expected " + std::to_string(p.n) + " ciphertexts, got " + std::to_string(cts.size()));
  if (party_index >= p.n) throw PvwError(PVW_ERR_INVALID_PARAMETERS, "Party index " + std::to_string(party_index) + " exceeds maximum " + std::to_string(p.n - 1));
  std::vector<uint32_t> slots(cts.size());
  for (size_t d = 0; d < cts.size(); d++) slots[d] = cts[d].slot;
  uint32_t pidx = (uint32_t)party_index;
  std::vector<uint64_t> out(cts.size());
  std::lock_guard<std::mutex> g(cts[0].ctx->mu);
  cts[0].ctx->check(pvw_decrypt_batch(cts[0].ctx->get(), (uint32_t)cts.size(), slots.data(), 1, &pidx, sk.secret_coeffs.data(), out.data(), PVW_IO_HOST));
  return out;
}


// ---- the crate's `serde` feature: bincode::serialize / bincode::deserialize of the polynomial-bearing types ----------
// (encryption.rs:298-354, crs.rs:228-295, public_key.rs:471-519; byte layout in pvw_b200.h).  The bit packing runs on the
// device where the data lives; only the wire bytes cross PCIe.
inline pvw_wire_layout wire_layout(const Context& ctx) {
  pvw_wire_layout w{};
  ctx.check(pvw_wire_layout_get(ctx.get(), &w));
  return w;
}
// bincode::serialize(&ciphertext)
inline std::vector<uint8_t> serialize(const PvwCiphertext& ct) {
  std::vector<uint8_t> out(wire_layout(*ct.ctx).ciphertext_bytes);
  std::lock_guard<std::mutex> g(ct.ctx->mu);
  ct.ctx->check(pvw_wire_ct_serialize(ct.ctx->get(), ct.slot, 1, out.data(), out.size(), PVW_IO_HOST));
  return out;
}
// bincode::deserialize::<PvwCiphertext>(&bytes) into store slot `slot` of the key's context; the embedded parameters must be
// the key's (the reference builds a fresh Arc<PvwParameters> from them instead)
inline PvwCiphertext deserialize_ciphertext(const std::vector<uint8_t>& bytes, const GlobalPublicKey& pk, uint32_t slot) {
  auto ctx = pk.context();
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->check(pvw_wire_ct_deserialize(ctx->get(), slot, 1, bytes.data(), bytes.size(), PVW_IO_HOST));
  return PvwCiphertext{pk.params, ctx, slot};
}
// bincode::serialize(&global_pk.crs) and the rows of global_pk.matrix ([u64 k][k records] each: one PublicKey body)
inline std::vector<uint8_t> serialize_crs(const GlobalPublicKey& pk) {
  auto ctx = pk.context();
  std::vector<uint8_t> out(wire_layout(*ctx).crs_bytes);
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->check(pvw_wire_crs_serialize(ctx->get(), out.data(), out.size(), PVW_IO_HOST));
  return out;
}
inline std::vector<uint8_t> serialize_public_key_rows(const GlobalPublicKey& pk, uint32_t row, uint32_t count) {
  auto ctx = pk.context();
  std::vector<uint8_t> out((size_t)count * wire_layout(*ctx).pk_row_bytes);
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->check(pvw_wire_pk_serialize_rows(ctx->get(), row, count, out.data(), PVW_IO_HOST));
  return out;
}
inline void deserialize_public_key_rows(GlobalPublicKey& pk, uint32_t row, uint32_t count, const std::vector<uint8_t>& bytes) {
  auto ctx = pk.context();
  if (bytes.size() < (size_t)count * wire_layout(*ctx).pk_row_bytes) throw PvwError(PVW_ERR_INSUFFICIENT_DATA, "public key rows: buffer too short");
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->check(pvw_wire_pk_deserialize_rows(ctx->get(), row, count, bytes.data(), PVW_IO_HOST));
}
// bincode::serialize(&*params)
inline std::vector<uint8_t> serialize(const PvwParameters& p) {
  std::vector<uint8_t> out(wire_layout(*p.probe).params_bytes);
  p.probe->check(pvw_wire_params(p.probe->get(), out.data(), out.size()));
  return out;
}

}  // namespace pvw
