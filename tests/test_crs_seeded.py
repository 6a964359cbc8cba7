"""Seeded CRS generation (src/params/crs.rs:45-90; reference tests tests/params.rs:87-174).
CPU: the library's host-side expansion against the Python restatement, plus known answers for the primitives whose
published vectors exist (ChaCha8, SHA-256).  The fhe-math / rand sampling rules are recalled: parity unpinned."""
import ctypes
import hashlib

import numpy as np
import pytest

import pvw_oracle as O


@pytest.fixture(scope="module")
def lib():
    import pvw_rs_b200
    return pvw_rs_b200._ffi.load()


def expand(lib, k, l, moduli, seed):
    mods = np.array(moduli, dtype=np.uint64)
    out = np.empty((k, k, len(moduli), l), dtype=np.uint64)
    assert lib.pvw_crs_expand_seed(k, l, len(moduli), mods.ctypes.data, bytes(seed), out.ctypes.data) == 0
    return out


def test_chacha8_known_answer():
    # ChaCha8, all-zero key and nonce, first block (the published "TC1" keystream of the reduced-round test vectors)
    ks = b"".join(x.to_bytes(4, "little") for x in O._chacha_block([0] * 8, 0, 8))
    assert ks.hex() == ("3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e"
                        "984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42")


@pytest.mark.parametrize("k,l,moduli", [(3, 8, O.TEST_MODULI), (2, 16, O.EX_MODULI), (2, 8, O.largest_ntt_primes(3))])
def test_library_expansion_equals_python_restatement(lib, k, l, moduli):
    seed = bytes(range(32))
    P = O.Params(1, k, l, moduli)
    want = np.array(O.crs_new_deterministic(P, seed), dtype=np.uint64)
    got = expand(lib, k, l, moduli, seed)
    assert got.shape == want.shape and (got == want).all()
    q = np.array(moduli, dtype=np.uint64).reshape(1, 1, -1, 1)
    assert (got < q).all()


def test_same_seed_same_crs_different_seed_different_crs(lib):
    # tests/params.rs:87-131
    a = expand(lib, 4, 8, O.TEST_MODULI, b"\x2a" * 32)
    b = expand(lib, 4, 8, O.TEST_MODULI, b"\x2a" * 32)
    c = expand(lib, 4, 8, O.TEST_MODULI, b"\x7b" * 32)
    assert (a == b).all() and (a != c).any()


def test_tags(lib):
    # tests/params.rs:133-174: same tag => same CRS, different tag => different CRS
    def seed(tag):
        s = (ctypes.c_uint8 * 32)()
        assert lib.pvw_crs_tag_to_seed(tag.encode(), s) == 0
        return bytes(s)
    assert seed("test_tag") == seed("test_tag") != seed("other_tag")
    assert seed("test_tag") == O.crs_seed_from_tag("test_tag") and seed("") == O.crs_seed_from_tag("")
    assert seed("x")[:8] * 4 == seed("x")                                  # the u64 cycled to 32 bytes (crs.rs:84-87)
    assert (expand(lib, 2, 8, O.EX_MODULI, seed("pvss-round-1")) == np.array(
        O.crs_new_deterministic(O.Params(1, 2, 8, O.EX_MODULI), O.crs_seed_from_tag("pvss-round-1")), dtype=np.uint64)).all()


def test_sha256_inside_the_expansion(lib):
    """the first polynomial of the matrix = ChaCha8(SHA-256(first element seed)) -> Uniform samples; pin the SHA-256 step"""
    seed = b"\x01" * 32
    master = O.chacha8_from_seed(seed)
    element_seed = bytes(master.next_u32() & 0xFF for _ in range(32))
    prng = O.chacha8_from_seed(hashlib.sha256(element_seed).digest())
    q = O.TEST_MODULI[0]
    first = O.uniform_u64_sample(prng, q)
    assert int(expand(lib, 1, 8, O.TEST_MODULI, seed)[0, 0, 0, 0]) == first


@pytest.mark.gpu
def test_api_level(lib):
    import pvw_rs_b200 as pvw
    params = pvw.PvwParametersBuilder().set_parties(3).set_dimension(4).set_l(8).set_moduli(O.TEST_MODULI).build()
    c1, c2 = pvw.PvwCrs.new_deterministic(params, b"\x2a" * 32), pvw.PvwCrs.new_deterministic(params, b"\x2a" * 32)
    assert (c1.matrix == c2.matrix).all() and c1.dimensions() == (4, 4)
    t1, t2 = pvw.PvwCrs.new_from_tag(params, "test_tag"), pvw.PvwCrs.new_from_tag(params, "other")
    assert (t1.matrix != t2.matrix).any()
    eng = params.new_engine()
    m = eng.crs_generate_from_tag("test_tag", want_matrix=True)
    assert (m == t1.matrix).all() and (eng.crs_download() == m).all()
