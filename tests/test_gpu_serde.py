"""GPU: the reference's serialisation tests (tests/serialization.rs, feature `serde`) restated against pvw_rs_b200.serde --
same parameters, same scenarios, same assertions -- plus byte equality with the CPU restatement (oracle/pvw_wire.py)."""
import numpy as np
import pytest

import pvw_oracle as O
import pvw_wire as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pvw():
    import pvw_rs_b200
    return pvw_rs_b200


def create_test_params(pvw):
    # tests/serialization.rs:10-20
    return (pvw.PvwParameters.builder().set_parties(3).set_dimension(2).set_l(8).set_moduli([0xffffee001, 0xffffc4001])
            .set_secret_variance(1.0).set_error_bounds_u32(50, 100).build_arc())


def oracle_params(p):
    return O.Params(p.n, p.k, p.l, p.moduli(), p.secret_variance, p.error_bound_1, p.error_bound_2, psi=p.psi)


def same_params(a, b):
    return (a.n, a.k, a.l, a.moduli(), a.secret_variance, a.error_bound_1, a.error_bound_2) == \
           (b.n, b.k, b.l, b.moduli(), b.secret_variance, b.error_bound_1, b.error_bound_2)


def full_key(pvw, params, rng):
    crs = pvw.PvwCrs.new(params, rng)
    gpk = pvw.GlobalPublicKey.new(crs)
    sks = [pvw.SecretKey.random(params, rng) for _ in range(params.n)]
    for i, sk in enumerate(sks):
        gpk.generate_and_add(i, sk, rng)
    return crs, gpk, sks


def test_pvw_parameters_serialization(pvw):
    # :22-40
    params = create_test_params(pvw)
    blob = pvw.serde.serialize(params)
    assert blob == W.params_to_bytes(oracle_params(params))
    rec = pvw.serde.deserialize(pvw.PvwParameters, blob)
    assert same_params(params, rec) and rec.delta == params.delta and rec.q_total() == params.q_total()


def test_secret_key_serialization(pvw):
    # :42-78
    params = create_test_params(pvw)
    sk = pvw.SecretKey.random(params, np.random.default_rng(1))
    blob = pvw.serde.serialize(sk)
    assert blob == W.secret_key_to_bytes(oracle_params(params), sk.coefficients().tolist())
    rec = pvw.serde.deserialize(pvw.SecretKey, blob)
    assert (rec.coefficients() == sk.coefficients()).all() and same_params(rec.params, params)


def test_public_key_serialization(pvw):
    # :80-130
    params = create_test_params(pvw)
    rng = np.random.default_rng(2)
    crs = pvw.PvwCrs.new(params, rng)
    pk = pvw.PublicKey.generate(pvw.SecretKey.random(params, rng), crs, rng)
    blob = pvw.serde.serialize(pk)
    assert blob == W.public_key_to_bytes(oracle_params(params), pk.key_polynomials.tolist())
    rec = pvw.serde.deserialize(pvw.PublicKey, blob)
    assert (rec.key_polynomials == pk.key_polynomials).all() and rec.dimension() == params.k and same_params(rec.params, params)


def test_pvw_crs_serialization(pvw):
    # :132-166
    params = create_test_params(pvw)
    crs = pvw.PvwCrs.new(params, np.random.default_rng(3))
    blob = pvw.serde.serialize(crs)
    assert blob == W.crs_to_bytes(oracle_params(params), crs.matrix.tolist())
    rec = pvw.serde.deserialize(pvw.PvwCrs, blob)
    assert rec.matrix.shape == crs.matrix.shape and (rec.matrix == crs.matrix).all() and same_params(rec.params, params)


def test_global_public_key_serialization(pvw):
    # :168-231: one party key added, the other rows still zero
    params = create_test_params(pvw)
    rng = np.random.default_rng(4)
    crs = pvw.PvwCrs.new(params, rng)
    gpk = pvw.GlobalPublicKey.new(crs)
    gpk.generate_and_add(0, pvw.SecretKey.random(params, rng), rng)
    blob = pvw.serde.serialize(gpk)
    P = oracle_params(params)
    assert blob == W.global_public_key_to_bytes(P, gpk.matrix.tolist(), crs.matrix.tolist(), 1, [])
    rec = pvw.serde.deserialize(pvw.GlobalPublicKey, blob)
    assert rec.matrix.shape == gpk.matrix.shape and (rec.matrix == gpk.matrix).all()
    assert (rec.crs().matrix == crs.matrix).all() and rec.num_keys == gpk.num_keys == 1 and same_params(rec.params, params)
    # with captured error polynomials (public_key.rs:304-329): party 1 only -> entry 0 is an empty Vec
    gpk.generate_and_add_with_errors(1, pvw.SecretKey.random(params, rng), rng)
    blob = pvw.serde.serialize(gpk)
    errs = [[], gpk.get_party_errors(1).tolist()]
    assert blob == W.global_public_key_to_bytes(P, gpk.matrix.tolist(), crs.matrix.tolist(), 2, errs)
    rec = pvw.serde.deserialize(pvw.GlobalPublicKey, blob)
    assert rec.num_keys == 2 and len(rec.get_all_errors()) == 2 and len(rec.get_party_errors(0)) == 0
    assert (rec.get_party_errors(1) == gpk.get_party_errors(1)).all() and (rec.matrix == gpk.matrix).all()


def test_ciphertext_serialization(pvw):
    # :233-295
    params = create_test_params(pvw)
    rng = np.random.default_rng(5)
    crs, gpk, sks = full_key(pvw, params, rng)
    ct = pvw.encrypt([i + 1 for i in range(params.n)], gpk)
    blob = pvw.serde.serialize(ct)
    assert blob == W.ciphertext_to_bytes(oracle_params(params), ct.c1.tolist(), ct.c2.tolist())
    rec = pvw.serde.deserialize(pvw.PvwCiphertext, blob, gpk)
    assert len(rec.c1) == len(ct.c1) and len(rec.c2) == len(ct.c2)
    assert (rec.c1 == ct.c1).all() and (rec.c2 == ct.c2).all() and same_params(rec.params, params)
    for i, sk in enumerate(sks):
        assert pvw.decrypt_party_value(rec, sk, i) == i + 1


def test_round_trip_consistency(pvw):
    # :297-317
    params = create_test_params(pvw)
    b1 = pvw.serde.serialize(params)
    r1 = pvw.serde.deserialize(pvw.PvwParameters, b1)
    b2 = pvw.serde.serialize(r1)
    r2 = pvw.serde.deserialize(pvw.PvwParameters, b2)
    assert b1 == b2 and same_params(params, r2)


def test_bincode_direct_usage(pvw):
    # :319-360
    params = create_test_params(pvw)
    crs, gpk, sks = full_key(pvw, params, np.random.default_rng(6))
    ct = pvw.encrypt([42] * params.n, gpk)
    rec = pvw.serde.deserialize(pvw.PvwCiphertext, pvw.serde.serialize(ct), gpk)
    assert len(rec) == len(ct) == params.n
    assert (rec.c1 == ct.c1).all() and (rec.c2 == ct.c2).all()
    rec.validate()
    # many ciphertexts in one device pass
    cts = pvw.encrypt_all_party_shares([[7 * d + p for p in range(params.n)] for d in range(params.n)], gpk)
    blobs = pvw.serde.serialize_ciphertexts(cts)
    P = oracle_params(params)
    for c, b in zip(cts, blobs):
        assert b.tobytes() == W.ciphertext_to_bytes(P, c.c1.tolist(), c.c2.tolist())


def test_serialization_deterministic(pvw):
    # :362-384
    params = create_test_params(pvw)
    sk = pvw.SecretKey.random(params, np.random.default_rng(7))
    b = [pvw.serde.serialize(sk) for _ in range(3)]
    assert b[0] == b[1] == b[2]
    assert (pvw.serde.deserialize(pvw.SecretKey, b[0]).coefficients() == sk.coefficients()).all()


def test_deserialize_rejects_foreign_and_damaged_blobs(pvw):
    params = create_test_params(pvw)
    crs, gpk, sks = full_key(pvw, params, np.random.default_rng(8))
    ct = pvw.encrypt([1, 2, 3], gpk)
    blob = bytearray(pvw.serde.serialize(ct))
    with pytest.raises(pvw.PvwError):
        pvw.serde.deserialize(pvw.PvwCiphertext, bytes(blob[:-3]), gpk)
    o = 8 + 8 + 6                                                           # Vec length, record length, Rq fields before the residues
    blob[o:o + 4] = b"\xff\xff\xff\xff"
    blob[o + 4] |= 0x0F                                                     # first residue of c1[0] = 2^36 - 1 >= q_0
    with pytest.raises(pvw.PvwError) as ei:
        pvw.serde.deserialize(pvw.PvwCiphertext, bytes(blob), gpk)
    assert ei.value.variant == "DeserializationError"
    other = (pvw.PvwParameters.builder().set_parties(3).set_dimension(2).set_l(8).set_moduli([0xffffee001, 0xffffc4001])
             .set_secret_variance(1.0).set_error_bounds_u32(50, 101).build_arc())
    crs2, gpk2, _ = full_key(pvw, other, np.random.default_rng(9))
    with pytest.raises(pvw.PvwError):
        pvw.serde.deserialize(pvw.PvwCiphertext, pvw.serde.serialize(ct), gpk2)     # embedded parameters differ
    with pytest.raises(pvw.PvwError):
        pvw.serde.deserialize(pvw.PvwParameters, pvw.serde.serialize(params)[:-1])
