"""Known-answer vectors from the real pvw crate (rust/pvw-dump) -- the pin that turns "parity unpinned" into "parity green".

No Rust toolchain exists in the build image and the reference holds no known-answer vectors of its own, so the directory
tests/golden/from_crate/ is empty until someone runs the dump program; every test here is skipped while it is.  With the files
in place the oracle (CPU) and the CUDA path (-m gpu) are compared with what fhe-math itself computed: psi and slot order
(NTT of X and of a ramp), Delta, encode_scalar, the seeded CRS (crs.rs:45-90), public-key rows (crs.rs:138-171 + public_key.rs:124-139),
one ciphertext under fixed r / e1 / e2 (encryption.rs:147-200), its decryption (decryption.rs:249-278) and the bincode bytes."""
import glob
import json
import os

import numpy as np
import pytest

import pvw_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "from_crate", "*.json")))
pytestmark = pytest.mark.skipif(not FILES, reason="tests/golden/from_crate/ holds no vectors (run rust/pvw-dump with cargo)")


def u64(rows):
    return np.array([[int(x) for x in r] for r in rows], dtype=np.uint64)


def load(path):
    d = json.load(open(path))
    moduli = [int(q) for q in d["moduli"]]
    ntt_x = u64(d["ntt_x"])
    psi = [int(ntt_x[j, 0]) for j in range(len(moduli))]        # slot 0 of NTT(X) is psi (slot i = psi^(2 brv(i) + 1))
    return d, moduli, psi


@pytest.mark.parametrize("path", FILES or ["none"])
def test_default_psi_and_slot_order_are_fhe_maths(path):
    d, moduli, psi = load(path)
    P = O.Params(d["n"], d["k"], d["l"], moduli, secret_variance=d["secret_variance"], error_bound_1=d["error_bound_1"], error_bound_2=d["error_bound_2"])
    assert list(P.psi) == psi, "the restated default psi selection (ChaCha8 seed 0 search) differs from fhe-math's"
    assert (np.array(P.ntt_forward(P.from_coefficients([0, 1] + [0] * (P.l - 2))), dtype=np.uint64) == u64(d["ntt_x"])).all()
    assert (np.array(P.ntt_forward(P.from_coefficients(d["ramp"])), dtype=np.uint64) == u64(d["ntt_ramp"])).all()
    assert P.delta == int(d["delta"]) and P.delta_power_l_minus_1 == int(d["delta_power_l_minus_1"])
    assert P.verify_correctness_condition() == d["correctness_condition"]
    assert (np.array(P.encode_scalar(12345), dtype=np.uint64) == u64(d["encode_scalar_12345"])).all()
    assert (np.array(P.encode_scalar(-7), dtype=np.uint64) == u64(d["encode_scalar_minus_7"])).all()


@pytest.mark.parametrize("path", FILES or ["none"])
def test_seeded_crs_keys_ciphertext_and_plaintexts(path):
    import c_oracle as CO
    d, moduli, psi = load(path)
    P = O.Params(d["n"], d["k"], d["l"], moduli, psi=psi, secret_variance=d["secret_variance"], error_bound_1=d["error_bound_1"], error_bound_2=d["error_bound_2"])
    A = O.crs_new_deterministic(P, bytes.fromhex(d["crs_seed_hex"]))
    assert (np.array(A[0][0], dtype=np.uint64) == u64(d["crs_a_0_0"])).all()
    assert (np.array(A[P.k - 1][P.k - 1], dtype=np.uint64) == u64(d["crs_a_last"])).all()
    At = O.crs_new_deterministic(P, O.crs_seed_from_tag(d["crs_tag"]))
    assert (np.array(At[0][0], dtype=np.uint64) == u64(d["crs_tag_a_0_0"])).all()
    co = CO.COracle(P)
    A_np = np.array(A, dtype=np.uint64)
    sk, ke = np.array(d["sk"], dtype=np.int64), np.array(d["key_error"], dtype=np.int64)
    B = co.keygen(A_np, sk, ke)
    assert (B[0] == np.stack([u64(p) for p in d["b_row_0"]])).all() and (B[-1] == np.stack([u64(p) for p in d["b_row_last"]])).all()
    m = np.array([[int(x) for x in d["m"]]], dtype=np.uint64)
    r, e1, e2 = (np.array(d[k_], dtype=np.int64)[None] for k_ in ("r", "e1", "e2"))
    c1, c2 = co.encrypt(A_np, B, m, r, e1, e2)
    assert (c1[0] == np.stack([u64(p) for p in d["c1"]])).all() and (c2[0] == np.stack([u64(p) for p in d["c2"]])).all()
    assert [int(x) for x in co.decrypt(sk, c1, c2)[:, 0]] == [int(x) for x in d["decrypted"]]


@pytest.mark.parametrize("path", FILES or ["none"])
def test_wire_bytes(path):
    import pvw_wire as W
    d, moduli, psi = load(path)
    P = O.Params(d["n"], d["k"], d["l"], moduli, psi=psi, secret_variance=d["secret_variance"], error_bound_1=d["error_bound_1"], error_bound_2=d["error_bound_2"])
    c1 = np.stack([u64(p) for p in d["c1"]])
    c2 = np.stack([u64(p) for p in d["c2"]])
    as_poly = lambda a: [[int(x) for x in row] for row in a]
    assert W.poly_to_bytes(P, as_poly(c1[0])).hex() == d["poly_to_bytes_hex_c1_0"]
    assert W.params_to_bytes(P).hex() == d["params_bincode_hex"]
    assert W.ciphertext_to_bytes(P, [as_poly(p) for p in c1], [as_poly(p) for p in c2]).hex() == d["ciphertext_bincode_hex"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES or ["none"])
def test_cuda_path_against_the_crate(path):
    import pvw_rs_b200 as pvw
    d, moduli, psi = load(path)
    eng = pvw.Engine(d["n"], d["k"], d["l"], moduli, secret_variance=d["secret_variance"], error_bound_1=d["error_bound_1"], error_bound_2=d["error_bound_2"])
    assert eng.psi == psi                                     # default psi derivation inside the library
    A = eng.crs_generate_deterministic(bytes.fromhex(d["crs_seed_hex"]), want_matrix=True)
    assert (A[0, 0] == u64(d["crs_a_0_0"])).all()
    sk, ke = np.array(d["sk"], dtype=np.int64), np.array(d["key_error"], dtype=np.int64)
    eng.keygen_batch(0, sk, ke)
    assert (eng.pk_download_rows(0, 1)[0] == np.stack([u64(p) for p in d["b_row_0"]])).all()
    eng.ct_reserve(1)
    m = np.array([[int(x) for x in d["m"]]], dtype=np.uint64)
    r, e1, e2 = (np.array(d[k_], dtype=np.int64)[None] for k_ in ("r", "e1", "e2"))
    eng.encrypt_batch(0, m, r, e1, e2)
    c1, c2 = eng.ct_download(0)
    assert (c1 == np.stack([u64(p) for p in d["c1"]])).all() and (c2 == np.stack([u64(p) for p in d["c2"]])).all()
    got = eng.decrypt_batch(np.arange(d["n"]), sk, D=1)
    assert [int(x) for x in got[:, 0]] == [int(x) for x in d["decrypted"]]
    assert eng.wire_ct_serialize(0, 1)[0].tobytes().hex() == d["ciphertext_bincode_hex"]
