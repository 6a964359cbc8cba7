"""CPU: the wire-format restatement (oracle/pvw_wire.py) against the behaviour the reference's own serialisation tests pin
(tests/serialization.rs): bincode round trips of every struct, byte determinism, structure of the blobs; plus hand-checked
known answers for the recalled third-party encodings (bit packing, protobuf framing, bincode framing)."""
import struct

import numpy as np
import pytest

import pvw_oracle as O
import pvw_wire as W
from _cases import System, params


def test_transcode_known_answers():
    # fhe-util transcode_to_bytes: LSB-first bit stream
    assert W.transcode_to_bytes([1, 2, 3, 4, 5, 6, 7, 0], 3) == bytes([0b11_010_001, 0b0_101_100_0, 0b000_111_11])
    assert W.transcode_to_bytes([0x1FF, 0, 0x155, 0x0AA, 1, 2, 3, 4], 9).hex() == \
        sum(v << (9 * i) for i, v in enumerate([0x1FF, 0, 0x155, 0x0AA, 1, 2, 3, 4])).to_bytes(9, "little").hex()
    vals = [(0x3FFFFFFFFFFFFFFF - 977 * i) for i in range(8)]
    b = W.transcode_to_bytes(vals, 62)
    assert len(b) == 62 and W.transcode_from_bytes(b, 62, 8) == vals
    assert W.nbits_of(0x3FFFFFFFFFFFFDC1) == 62 and W.nbits_of(0xFFFFC4001) == 36 and W.nbits_of(0x1FFFFE0001) == 37
    assert W.nbits_of(17) == 5 and W.nbits_of(257) == 9                       # p - 1 a power of two


def test_varint_and_rq_framing():
    assert W.varint(8) == b"\x08" and W.varint(127) == b"\x7f" and W.varint(128) == b"\x80\x01" and W.varint(1054) == b"\x9e\x08"
    P = params("EX")
    zero = [[0] * P.l for _ in P.moduli]
    b = W.poly_to_bytes(P, zero)
    packed = (36 + 37) * P.l // 8
    assert b[:6] == bytes([0x08, 0x02, 0x10, 0x08, 0x1A, packed]) and len(b) == 6 + packed
    assert W.poly_record_bytes(P) == 8 + len(b)
    rows, rep = W.poly_from_bytes(P, b)
    assert rows == zero and rep == W.REP_NTT
    # power-basis polynomials carry representation 1
    assert W.poly_to_bytes(P, zero, W.REP_POWERBASIS)[:2] == b"\x08\x01"


def test_params_blob_layout():
    # impl Serialize for PvwParameters, parameters.rs:606-623 ; tests/serialization.rs:22-40
    P = params("EX")
    b = W.params_to_bytes(P)
    assert b[:24] == struct.pack("<QQQ", 7, 32, 8)
    assert b[24:32] == struct.pack("<Q", 2) and b[32:48] == struct.pack("<QQ", *O.EX_MODULI)
    assert b[48:52] == struct.pack("<f", 0.5)
    assert b[52:] == struct.pack("<Q", 2) + b"50" + struct.pack("<Q", 2) + b"50"
    R = W.params_from_bytes(b, psi=P.psi)
    assert W.same_params(P, R) and R.delta == P.delta and R.Q == P.Q
    # tests/serialization.rs:297-317: serialize(deserialize(serialize(x))) == serialize(x)
    assert W.params_to_bytes(R) == b
    # trailing bytes are tolerated by bincode::deserialize; truncation is not
    W.params_from_bytes(b + b"\x00", psi=P.psi)
    with pytest.raises(O.PvwError):
        W.params_from_bytes(b[:-1], psi=P.psi)
    # the deserialiser re-runs the builder (parameters.rs:652-661): bad parameters are rejected
    bad = bytearray(b)
    bad[16:24] = struct.pack("<Q", 7)                                        # l = 7
    with pytest.raises(O.PvwError):
        W.params_from_bytes(bytes(bad), psi=P.psi)


@pytest.mark.parametrize("name", ["EX", "T16", "RAG"])
def test_struct_round_trips(name):
    # tests/serialization.rs:42-295: every struct survives bincode, polynomials compared through to_bytes()
    P = params(name)
    S = System(P, 2)
    c1, c2 = S.encrypt()
    A, B = S.A.tolist(), S.B.tolist()
    # secret key (:42-78)
    b = W.secret_key_to_bytes(P, S.sk[0].tolist())
    P2, coeffs = W.secret_key_from_bytes(b, psi=P.psi)
    assert W.same_params(P, P2) and coeffs == S.sk[0].tolist()
    assert W.secret_key_to_bytes(P, S.sk[0].tolist()) == b                   # determinism (:362-384)
    # public key (:80-130)
    b = W.public_key_to_bytes(P, B[1])
    P2, polys = W.public_key_from_bytes(b, psi=P.psi)
    assert polys == B[1] and W.same_params(P, P2)
    assert len(b) == 8 + P.k * W.poly_record_bytes(P) + len(W.params_to_bytes(P))
    # CRS (:132-166)
    b = W.crs_to_bytes(P, A)
    P2, A2 = W.crs_from_bytes(b, psi=P.psi)
    assert A2 == A
    assert len(b) == 8 + P.k * (8 + P.k * W.poly_record_bytes(P)) + len(W.params_to_bytes(P))
    # global public key (:168-231), error polynomials in NTT form for the first two parties only
    errs = [[P.ntt_forward(P.from_coefficients(row)) for row in S.ke[i].tolist()] for i in range(2)]
    b = W.global_public_key_to_bytes(P, B, A, P.n, errs)
    P2, B2, A2, nk, errs2 = W.global_public_key_from_bytes(b, psi=P.psi)
    assert B2 == B and A2 == A and nk == P.n and errs2 == errs
    # ciphertext (:233-295, :319-360)
    b = W.ciphertext_to_bytes(P, c1[0].tolist(), c2[0].tolist())
    P2, d1, d2 = W.ciphertext_from_bytes(b, psi=P.psi)
    assert d1 == c1[0].tolist() and d2 == c2[0].tolist() and W.same_params(P, P2)
    assert len(b) == 16 + (P.k + P.n) * W.poly_record_bytes(P) + len(W.params_to_bytes(P))
    assert W.ciphertext_to_bytes(P2, d1, d2) == b


def test_poly_from_bytes_rejects_malformed():
    P = params("EX")
    S = System(P, 1)
    good = W.poly_to_bytes(P, S.A[0][0].tolist())
    for bad in (good[:-1], good + b"\x00\x00", b"\x08\x03" + good[2:], good[:3] + b"\x10" + good[4:]):
        with pytest.raises(O.PvwError):
            W.poly_from_bytes(P, bad)
    # a residue >= q_0: all-ones in the first 36-bit field
    bad = bytearray(good)
    bad[6:10] = b"\xff\xff\xff\xff"
    bad[10] |= 0x0F
    with pytest.raises(O.PvwError):
        W.poly_from_bytes(P, bytes(bad))


def test_golden_wire_fixture():
    """tests/golden/EX_wire.npz (made by tests/golden/make_golden.py) pins this restatement against accidental change"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "EX_wire.npz"))
    P = params("EX")
    S = System(P, 2)
    c1, c2 = S.encrypt()
    assert W.ciphertext_to_bytes(P, c1[1].tolist(), c2[1].tolist()) == g["ct1"].tobytes()
    assert W.params_to_bytes(P) == g["params"].tobytes()
    assert W.public_key_to_bytes(P, S.B[3].tolist()) == g["pk3"].tobytes()
