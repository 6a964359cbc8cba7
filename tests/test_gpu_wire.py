"""GPU: the device serialisers / parsers of the wire format (csrc/wire.cu, through the C ABI) against oracle/pvw_wire.py --
byte-exact blobs, exact round trips, and the rejections the reference's deserialisers make (tests/serialization.rs)."""
import os
import struct

import numpy as np
import pytest

import pvw_oracle as O
import pvw_wire as W
from _cases import P128_MODULI, System, engine_kwargs, params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pvw():
    import pvw_rs_b200
    return pvw_rs_b200


def loaded(pvw, S, D, **over):
    P = S.P
    eng = pvw.Engine(**engine_kwargs(P, **over))
    eng.crs_upload(S.A)
    eng.pk_upload_rows(eng.row0, S.B[eng.row0:eng.row0 + eng.nrows])
    eng.ct_reserve(D)
    rows = slice(eng.row0, eng.row0 + eng.nrows)
    eng.encrypt_batch(0, S.m[:, rows], S.r, S.e1, S.e2[:, rows])
    return eng


@pytest.mark.parametrize("name,D", [("EX", 3), ("T16", 2), ("RAG", 3), ("L32", 2), ("VDs", 2), ("P128s", 2), ("P256s", 1)])
def test_ciphertext_blobs_are_byte_exact(pvw, name, D):
    P = params(name)
    S = System(P, D)
    eng = loaded(pvw, S, D)
    c1, c2 = S.encrypt()
    lay = eng.wire_layout
    assert lay.record_bytes == W.poly_record_bytes(P) and eng.wire_params() == W.params_to_bytes(P)
    blobs = eng.wire_ct_serialize(0, D)
    assert blobs.shape == (D, lay.ciphertext_bytes)
    for d in range(D):
        assert blobs[d].tobytes() == W.ciphertext_to_bytes(P, c1[d].tolist(), c2[d].tolist()), (name, d)
    # parse into fresh slots of a second context and compare the residues (tests/serialization.rs:233-295)
    eng2 = pvw.Engine(**engine_kwargs(P))
    eng2.ct_reserve(D + 1)
    eng2.wire_ct_deserialize(1, D, blobs)
    for d in range(D):
        g1, g2 = eng2.ct_download(1 + d)
        assert (g1 == c1[d]).all() and (g2 == c2[d]).all()
    # ... and the parsed ciphertexts decrypt (test_bincode_direct_usage, :319-360)
    out = eng2.decrypt_batch(np.arange(P.n, dtype=np.uint32), S.sk, dealer_slots=np.arange(1, D + 1, dtype=np.uint32))
    assert (out == S.m.T).all()
    # serialize(deserialize(bytes)) == bytes (:297-317); a wider stride leaves the gap bytes alone
    wide = np.full((D, lay.ciphertext_bytes + 13), 0xA5, dtype=np.uint8)
    eng2.wire_ct_serialize(1, D, out=wide, stride=wide.shape[1])
    assert (wide[:, :lay.ciphertext_bytes] == blobs).all() and (wide[:, lay.ciphertext_bytes:] == 0xA5).all()


def test_device_resident_blobs(pvw):
    """PVW_IO_DEVICE: blobs are produced into / parsed from a CUDA byte tensor without touching the host, at odd offsets"""
    import torch
    P = params("P128s")
    D = 5
    S = System(P, D)
    eng = loaded(pvw, S, D)
    lay = eng.wire_layout
    ref = eng.wire_ct_serialize(0, D)
    stride = lay.ciphertext_bytes + 3                                       # misaligned blobs
    buf = torch.zeros(1 + D * stride, dtype=torch.uint8, device="cuda")
    eng.wire_ct_serialize(0, D, out=buf[1:], stride=stride)
    got = buf[1:].cpu().numpy().reshape(D, stride)
    assert (got[:, :lay.ciphertext_bytes] == ref).all() and (got[:, lay.ciphertext_bytes:] == 0).all()
    eng.ct_reserve(D)                                                       # wipes the store
    eng.wire_ct_deserialize(0, D, buf[1:], stride=stride)
    c1, c2 = S.encrypt()
    for d in range(D):
        g1, g2 = eng.ct_download(d)
        assert (g1 == c1[d]).all() and (g2 == c2[d]).all()


@pytest.mark.parametrize("name", ["EX", "RAG", "P128s"])
def test_public_key_and_crs_blobs(pvw, name):
    P = params(name)
    S = System(P, 1)
    eng = loaded(pvw, S, 1)
    lay = eng.wire_layout
    tail = W.params_to_bytes(P)
    rows = eng.wire_pk_serialize_rows(0, P.n).tobytes()
    for i in range(P.n):
        row = rows[i * lay.pk_row_bytes:(i + 1) * lay.pk_row_bytes]
        assert row + tail == W.public_key_to_bytes(P, S.B[i].tolist())      # PublicKey, public_key.rs:471-487
    crs = eng.wire_crs_serialize()
    assert crs == W.crs_to_bytes(P, S.A.tolist())                           # PvwCrs, crs.rs:228-249
    # GlobalPublicKey (public_key.rs:522-552) assembled from the pieces
    errs = eng.ntt_forward_small(S.ke[:2].reshape(-1, P.l)).reshape(2, P.k, P.L, P.l)
    err_bytes = struct.pack("<Q", 2) + b"".join(struct.pack("<Q", P.k) + eng.wire_polys_serialize(e) for e in errs)
    gpk = struct.pack("<Q", P.n) + rows + crs + struct.pack("<Q", P.n) + tail + err_bytes
    assert gpk == W.global_public_key_to_bytes(P, S.B.tolist(), S.A.tolist(), P.n, errs.tolist())
    # round trip into an empty context
    eng2 = pvw.Engine(**engine_kwargs(P))
    eng2.wire_crs_deserialize(crs)
    eng2.wire_pk_deserialize_rows(0, P.n, rows)
    assert (eng2.crs_download() == S.A).all() and (eng2.pk_download_rows(0, P.n) == S.B).all() and eng2.num_keys == P.n
    assert (eng2.wire_polys_deserialize(eng.wire_polys_serialize(errs[1]), P.k) == errs[1]).all()


def test_sharded_contexts_write_disjoint_byte_ranges(pvw):
    """row-sharded contexts: each writes the envelope, c1 and its own parties' c2 records; the union is the blob"""
    P = params("RAG")
    D = 2
    S = System(P, D)
    c1, c2 = S.encrypt()
    plan = [pvw.sharding.ShardPlan(P.n, 2, r) for r in range(2)]
    engs = [loaded(pvw, S, D, row0=p.row0, nrows=p.nrows) for p in plan]
    lay = engs[0].wire_layout
    parts = [e.wire_ct_serialize(0, D) for e in engs]
    for d in range(D):
        want = np.frombuffer(W.ciphertext_to_bytes(P, c1[d].tolist(), c2[d].tolist()), dtype=np.uint8)
        merged = parts[0][d].copy()
        lo = lay.ct_c2_offset + 8 + plan[1].row0 * lay.record_bytes
        hi = lo + plan[1].nrows * lay.record_bytes
        assert (parts[0][d][lo:hi] == 0).all()
        merged[lo:hi] = parts[1][d][lo:hi]
        assert (merged == want).all()
        # each shard parses the full blob, keeping only its rows
    fresh = [pvw.Engine(**engine_kwargs(P, row0=p.row0, nrows=p.nrows)) for p in plan]
    full = np.stack([np.frombuffer(W.ciphertext_to_bytes(P, c1[d].tolist(), c2[d].tolist()), dtype=np.uint8) for d in range(D)])
    for e, p in zip(fresh, plan):
        e.ct_reserve(D)
        e.wire_ct_deserialize(0, D, full)
        g1, g2 = e.ct_download(1)
        assert (g1 == c1[1]).all() and (g2 == c2[1][p.row0:p.row0 + p.nrows]).all()


def test_malformed_blobs_are_rejected_and_leave_the_store_untouched(pvw):
    P = params("EX")
    S = System(P, 1)
    eng = loaded(pvw, S, 1)
    lay = eng.wire_layout
    good = eng.wire_ct_serialize(0, 1)[0]
    before = eng.ct_download(0)

    def rejected(blob, variant="DeserializationError"):
        with pytest.raises(pvw.PvwError) as ei:
            eng.wire_ct_deserialize(0, 1, blob)
        assert ei.value.variant == variant, ei.value
        after = eng.ct_download(0)
        assert (after[0] == before[0]).all() and (after[1] == before[1]).all()

    rejected(good[:-1], "InsufficientData")                                  # truncated
    bad = good.copy(); bad[0] ^= 1; rejected(bad)                            # c1 count != k
    bad = good.copy(); bad[lay.ct_c2_offset] ^= 2; rejected(bad)             # c2 count != n
    bad = good.copy(); bad[8] ^= 1; rejected(bad)                            # record length prefix
    bad = good.copy(); bad[8 + 8 + 1] = 1; rejected(bad)                     # representation PowerBasis, not NTT
    bad = good.copy(); bad[8 + 8 + 3] = 16; rejected(bad)                    # degree
    bad = good.copy(); bad[lay.ct_params_offset] ^= 1; rejected(bad)         # embedded n differs
    bad = good.copy(); bad[-1] = ord("1"); rejected(bad)                     # embedded error bound differs
    # a residue >= q_0 in the last c2 record (first 36-bit field all ones)
    o = lay.ct_c2_offset + 8 + (P.n - 1) * lay.record_bytes + 8 + 6
    bad = good.copy(); bad[o:o + 4] = 0xFF; bad[o + 4] |= 0x0F; rejected(bad)
    with pytest.raises(O.PvwError):
        W.ciphertext_from_bytes(bad.tobytes(), psi=P.psi)                    # the oracle rejects the same bytes
    eng.wire_ct_deserialize(0, 1, good)                                      # and the good blob still loads
    # CRS / public-key parsers
    crs = np.frombuffer(eng.wire_crs_serialize(), dtype=np.uint8)
    bad = crs.copy(); bad[16 + 8 + 5] ^= 0x40
    with pytest.raises(pvw.PvwError):
        eng.wire_crs_deserialize(bad.tobytes())
    assert (eng.crs_download() == S.A).all()
    rows = eng.wire_pk_serialize_rows(2, 1)
    with pytest.raises(pvw.PvwError) as ei:
        eng.wire_pk_deserialize_rows(P.n, 1, rows)
    assert ei.value.variant == "IndexOutOfBounds"


def test_golden_wire_fixture(pvw):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "EX_wire.npz"))
    P = params("EX")
    S = System(P, 2)
    eng = loaded(pvw, S, 2)
    assert eng.wire_ct_serialize(1, 1)[0].tobytes() == g["ct1"].tobytes()
    assert eng.wire_crs_serialize() == g["crs"].tobytes()
    assert eng.wire_pk_serialize_rows(3, 1).tobytes() + eng.wire_params() == g["pk3"].tobytes()


def test_small_modulus_bit_widths(pvw):
    """records whose residues are narrower than a byte boundary pattern: 5-, 9- and 13-bit primes (p - 1 a power of two
    for the first two), several residues per byte"""
    mods = [17, 257, 7681, 12289, 65537]
    P = O.Params(2, 3, 8, mods, error_bound_1=1, error_bound_2=1)
    eng = pvw.Engine(**engine_kwargs(P))
    rng = np.random.default_rng(3)
    polys = np.stack([[rng.integers(0, q, size=P.l, dtype=np.uint64) for q in mods] for _ in range(11)])
    polys[0] = np.array([[q - 1] * P.l for q in mods], dtype=np.uint64)
    b = eng.wire_polys_serialize(polys)
    rec = W.poly_record_bytes(P)
    for i in range(len(polys)):
        assert b[i * rec:(i + 1) * rec] == struct.pack("<Q", rec - 8) + W.poly_to_bytes(P, polys[i].tolist())
    assert (eng.wire_polys_deserialize(b, len(polys)) == polys).all()
