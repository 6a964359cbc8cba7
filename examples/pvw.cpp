// examples/pvw.cpp -- the reference's `cargo run --example pvw` scenario (examples/pvw.rs) on the C++ host mirror:
// multi-party keygen, share distribution with encrypt_all_party_shares, per-party decrypt_party_shares, verification.
// Also runs the error paths of tests/crypto.rs:181-207.  Exit code 0 iff every share is recovered.
//   usage: pvw_example [n] [k] [l]
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "pvw_b200.hpp"

using namespace pvw;
using clk = std::chrono::steady_clock;

template <class F>
static bool fails_with(const char* variant, F&& f) {
  try { f(); } catch (const PvwError& e) { return e.variant() == variant; }
  return false;
}

int main(int argc, char** argv) {
  const uint32_t num_parties = argc > 1 ? atoi(argv[1]) : 7;        // examples/pvw.rs:28-32
  const uint32_t dimension = argc > 2 ? atoi(argv[2]) : 32;
  const uint32_t ring_degree = argc > 3 ? atoi(argv[3]) : 8;
  const std::vector<uint64_t> moduli = {0xffffc4001ull, 0x1ffffe0001ull};
  try {
    auto params = PvwParametersBuilder().set_parties(num_parties).set_dimension(dimension).set_l(ring_degree).set_moduli(moduli)
                      .set_secret_variance(0.5f).set_error_bounds_u32(50, 50) /* suggest_error_bounds for this set */.build_arc();
    printf("PVW parameters: n=%u t=%u k=%u l=%u, %zu moduli, delta=%llu, correctness condition: %s\n", params->n, params->t, params->k, params->l,
           params->L(), (unsigned long long)params->delta[0], params->verify_correctness_condition() ? "ok" : "VIOLATED");
    Rng crs_rng, rng;   // OS-seeded CSPRNGs (the reference: thread_rng()); one for the public CRS, one for secrets
    auto t0 = clk::now();
    PvwCrs crs = PvwCrs::new_random(params, crs_rng);
    GlobalPublicKey global_pk(crs);
    std::vector<Party> parties;
    for (uint32_t i = 0; i < num_parties; i++) {
      parties.push_back(Party::make(i, params, rng));
      global_pk.generate_and_add_party(parties.back(), rng);        // examples/pvw.rs:88-92
    }
    auto t1 = clk::now();
    std::vector<std::vector<uint64_t>> all(num_parties, std::vector<uint64_t>(num_parties));
    for (uint32_t d = 0; d < num_parties; d++)
      for (uint32_t j = 1; j <= num_parties; j++) all[d][j - 1] = (uint64_t)d * 1000 + j;   // examples/pvw.rs:98-100
    auto cts = encrypt_all_party_shares(all, global_pk, rng);       // examples/pvw.rs:130-132
    auto t2 = clk::now();
    size_t ok = 0;
    for (uint32_t p = 0; p < num_parties; p++) {                    // examples/pvw.rs:135-152
      auto got = decrypt_party_shares(cts, parties[p].secret_key(), p);
      for (uint32_t d = 0; d < num_parties; d++) ok += got[d] == all[d][p];
    }
    auto t3 = clk::now();
    auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    printf("keygen %.2f ms, encrypt_all_party_shares %.2f ms, decrypt %.2f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3));
    printf("recovered %zu / %u shares (%.1f %%)\n", ok, num_parties * num_parties, 100.0 * ok / (num_parties * num_parties));
    // single-value path and shapes (tests/crypto.rs:91-149)
    bool good = ok == (size_t)num_parties * num_parties;
    good &= decrypt_party_value(cts[num_parties - 1], parties[0].secret_key(), 0) == all[num_parties - 1][0];
    good &= cts[0].c1().size() == (size_t)params->k * params->poly_words() && cts[0].c2().size() == (size_t)params->n * params->poly_words();
    // error paths (tests/crypto.rs:181-207, decryption.rs:286-309, encryption.rs:117)
    std::vector<uint64_t> short_vec(num_parties - 1, 1);
    good &= fails_with("InvalidParameters", [&] { encrypt(short_vec, global_pk, rng); });
    good &= fails_with("InvalidParameters", [&] { encrypt_party_shares(all[0], num_parties, global_pk, rng); });
    good &= fails_with("InvalidParameters", [&] { decrypt_party_shares({}, parties[0].secret_key(), 0); });
    good &= fails_with("InvalidParameters", [&] { decrypt_party_shares({cts[0]}, parties[0].secret_key(), 0); });
    good &= fails_with("InvalidParameters", [&] { decrypt_party_shares(cts, parties[0].secret_key(), num_parties); });
    {
      GlobalPublicKey partial(crs);
      partial.generate_and_add_party(parties[0], rng);
      good &= !partial.is_full();
      good &= fails_with("InvalidParameters", [&] { encrypt(all[0], partial, rng); });
    }
    {
      // tests/serialization.rs:233-295, :319-360: a ciphertext survives bincode byte for byte and still decrypts; a second
      // key holder rebuilt from the serialised public-key rows and CRS produces identical bytes
      auto bytes = serialize(cts[1]);
      auto back = deserialize_ciphertext(bytes, global_pk, 0);          // overwrites slot 0 with dealer 1's ciphertext
      good &= serialize(back) == bytes && back.c1() == cts[1].c1() && back.c2() == cts[1].c2();
      good &= decrypt_party_value(back, parties[2].secret_key(), 2) == all[1][2];
      GlobalPublicKey twin(crs);
      deserialize_public_key_rows(twin, 0, num_parties, serialize_public_key_rows(global_pk, 0, num_parties));
      good &= twin.is_full() && serialize_crs(twin) == serialize_crs(global_pk);
      good &= twin.get_public_key(3) == global_pk.get_public_key(3);
      bytes[bytes.size() - 1] ^= 1;                                      // embedded error bound changed
      good &= fails_with("DeserializationError", [&] { deserialize_ciphertext(bytes, global_pk, 0); });
      bytes.resize(bytes.size() - 10);
      good &= fails_with("InsufficientData", [&] { deserialize_ciphertext(bytes, global_pk, 0); });
      printf("serialised ciphertext: %zu bytes (raw residues %zu)\n", bytes.size() + 10, (size_t)(params->k + params->n) * params->poly_words() * 8);
    }
    good &= fails_with("InvalidParameters", [&] { PvwParametersBuilder().set_parties(3).set_dimension(4).set_l(12).set_moduli(moduli).build_arc(); });
    printf("%s\n", good ? "ALL CHECKS PASSED" : "CHECK FAILED");
    return good ? 0 : 1;
  } catch (const PvwError& e) {
    printf("PvwError: %s\n", e.what());
    return 2;
  }
}
