"""The oracle against the reference's own behavioural pins and against itself
(exact-Python vs C restatement, NTT path vs psi-independent schoolbook ring product)."""
import random

import numpy as np
import pytest

import c_oracle as CO
import pvw_oracle as O

CONFIGS = {
    "EX": dict(n=7, k=32, l=8, moduli=O.EX_MODULI, secret_variance=0.5, error_bound_1=50, error_bound_2=50),
    "T16": dict(n=10, k=4, l=16, moduli=O.TEST_MODULI, secret_variance=0.5, error_bound_1=50, error_bound_2=50),
    "VDs": dict(n=3, k=5, l=8, moduli=O.VD_MODULI, secret_variance=10.0, error_bound_1=1, error_bound_2=1172385),
    "L32": dict(n=3, k=3, l=32, moduli=O.largest_ntt_primes(5), secret_variance=1.0, error_bound_1=100, error_bound_2=200),
}


def make(name):
    return O.Params(**CONFIGS[name])


def test_parameter_sets_match_survey():
    P = make("EX")
    assert P.delta == 558 and P.Q.bit_length() == 73
    P3 = O.Params(3, 4, 8, O.TEST_MODULI)
    assert P3.delta == 12633 and P3.Q.bit_length() == 109
    pr = O.largest_ntt_primes(17)
    assert pr[0] == 0x3FFFFFFFFFFFFDC1 and pr[-1] == 0x3FFFFFFFFFFFBB81
    P128 = O.Params(4, 2, 8, pr)
    assert P128.Q.bit_length() == 1054 and P128.delta.bit_length() == 132
    assert P128.delta_power_l_minus_1.bit_length() == 923
    VD = O.Params(5, 4, 8, O.VD_MODULI, 10.0, 1, 1172385)
    assert VD.delta == 189812531 and VD.Q.bit_length() == 221
    assert O.Params.suggest_error_bounds(7, 32, 8, O.EX_MODULI, 0.5) == (50, 50)
    assert O.Params.suggest_error_bounds(10, 4, 16, O.TEST_MODULI, 0.5) == (50, 50)
    assert P.t == 3                                                            # parameters.rs:169


def test_builder_rejections():
    # parameters.rs:131-144 ; tests/keys.rs:540-576 (k=0 rejected)
    for kw in (dict(n=0, k=4, l=8), dict(n=3, k=0, l=8), dict(n=3, k=4, l=4), dict(n=3, k=4, l=12)):
        with pytest.raises(O.PvwError):
            O.Params(moduli=O.TEST_MODULI, **kw)
    with pytest.raises(O.PvwError):
        O.Params(3, 4, 8, [0xFFFFEE001, 0xFFFFEE001])
    with pytest.raises(O.PvwError):
        O.Params(3, 4, 8, O.TEST_MODULI, error_bound_1=0)


def test_rounding_division_rule():
    # tests/crypto.rs:307-330
    for dividend, divisor, expected in [(7, 3, 2), (8, 3, 3), (-7, 3, -2), (-8, 3, -3)]:
        tw = 2 * dividend
        q = O.tdiv(tw - divisor, 2 * divisor) if dividend < 0 else O.tdiv(tw + divisor, 2 * divisor)
        assert q == expected
    assert O.trem(-7, 3) == -1 and O.trem(7, -3) == 1 and O.tdiv(-7, 2) == -3


@pytest.mark.parametrize("name", list(CONFIGS))
def test_gadget_structure(name):
    # tests/crypto.rs:17-44,151-158 ; tests/params.rs:637-674
    P = make(name)
    coeffs = P.lift(P.ntt_backward(P.gadget_polynomial()))
    assert coeffs == [P.delta ** i for i in range(P.l)]
    P.encode_scalar(42)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_bigints_to_poly_semantics(name):
    # tests/params.rs:484-635,732-767
    P = make(name)
    assert P.lift(P.bigints_to_poly([0] * P.l)) == [0] * P.l
    small = list(range(1, P.l + 1))
    assert P.lift(P.bigints_to_poly(small)) == small
    assert P.bigints_to_poly(small) == P.from_coefficients(small)
    big = [P.delta * (i + 1) for i in range(P.l)]
    assert P.lift(P.bigints_to_poly(big)) == [b % P.Q for b in big]
    neg = [-(i + 1) * 1000003 for i in range(P.l)]
    assert P.lift(P.bigints_to_poly(neg)) == [x % P.Q for x in neg]
    mixed = [P.delta // 2, 1, -1, 10 ** 6, -(10 ** 6), 0, P.Q - 1, -(P.Q - 1)] + [0] * (P.l - 8)
    assert P.lift(P.bigints_to_poly(mixed)) == [x % P.Q for x in mixed]
    with pytest.raises(O.PvwError):
        P.bigints_to_poly([1] * (P.l - 1))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_ntt_is_negacyclic_ring_isomorphism(name):
    """psi-independent algebra: slot-wise product in the NTT domain == schoolbook product mod (X^l + 1, Q)."""
    P = make(name)
    rnd = random.Random(5)
    a = [rnd.randrange(P.Q) for _ in range(P.l)]
    b = [rnd.randrange(P.Q) for _ in range(P.l)]
    prod = [0] * P.l
    for i in range(P.l):
        for j in range(P.l):
            t = i + j
            if t < P.l:
                prod[t] = (prod[t] + a[i] * b[j]) % P.Q
            else:
                prod[t - P.l] = (prod[t - P.l] - a[i] * b[j]) % P.Q
    ah, bh = P.ntt_forward(P.bigints_to_poly(a)), P.ntt_forward(P.bigints_to_poly(b))
    assert P.lift(P.ntt_backward(P.mul(ah, bh))) == prod
    assert P.ntt_backward(ah) == P.bigints_to_poly(a)
    # slot i holds the evaluation at psi^(2*brv(i)+1)  
    lg = P.l.bit_length() - 1
    for j, q in enumerate(P.moduli):
        for i in (0, 1, P.l - 1):
            x = pow(P.psi[j], 2 * O.brv(i, lg) + 1, q)
            assert ah[j][i] == sum(c * pow(x, t, q) for t, c in enumerate(a)) % q
    # constants map to all-equal slots
    c = P.ntt_forward(P.bigints_to_poly([12345] + [0] * (P.l - 1)))
    assert all(row == [12345 % q] * P.l for row, q in zip(c, P.moduli))


def test_default_psi_is_primitive_and_deterministic():
    for q in O.TEST_MODULI + O.VD_MODULI + O.largest_ntt_primes(3):
        for l in (8, 16, 32):
            if (q - 1) % (2 * l):
                continue
            p = O.fhe_math_default_psi(q, l)
            assert pow(p, l, q) == q - 1
            assert p == O.fhe_math_default_psi(q, l)


def _system(P, D, msg_mode="example"):
    A = O.synth_crs(P)
    sk = O.synth_small(P, O.TAG_SK, P.n, P.k, "cbd")
    ke = O.synth_small(P, O.TAG_KE, P.n, P.k, "uniform", P.error_bound_1)
    B = [O.keygen(P, A, sk[p], ke[p]) for p in range(P.n)]
    m = O.synth_messages(P, D, msg_mode)
    r = O.synth_small(P, O.TAG_R, D, P.k, "cbd")
    e1 = O.synth_small(P, O.TAG_E1, D, P.k, "uniform", P.error_bound_1)
    e2 = O.synth_small(P, O.TAG_E2, D, P.n, "uniform", P.error_bound_2)
    return A, sk, ke, B, m, r, e1, e2


@pytest.mark.parametrize("name", ["EX", "T16", "VDs"])
def test_end_to_end_recovery_and_shapes(name):
    # tests/crypto.rs:91-149 (shapes) and :236-305 (>= 95 % recovered; here noise is within bounds => 100 %)
    P = make(name)
    D = P.n
    A, sk, ke, B, m, r, e1, e2 = _system(P, D)
    cts = [O.encrypt_explicit(P, A, B, m[d], r[d], e1[d], e2[d], num_keys=P.n) for d in range(D)]
    for c1, c2 in cts:
        assert len(c1) == P.k and len(c2) == P.n
    for p in range(P.n):
        got = O.decrypt_party_shares(P, cts, sk[p], p)
        assert got == [m[d][p] for d in range(D)]
    # literal restatement of decode_scalar_pvw_rns agrees with the scalar form
    assert O.decrypt_party_value(P, cts[0][0], cts[0][1], sk[1], 1, literal=True) == m[0][1]


def test_error_paths():
    # tests/crypto.rs:181-207 ; decryption.rs:286-309
    P = make("EX")
    A, sk, ke, B, m, r, e1, e2 = _system(P, 1)
    with pytest.raises(O.PvwError):
        O.encrypt_explicit(P, A, B, m[0][:-1], r[0], e1[0], e2[0])
    with pytest.raises(O.PvwError):
        O.encrypt_explicit(P, A, B, m[0] + [1], r[0], e1[0], e2[0])
    with pytest.raises(O.PvwError):
        O.encrypt_explicit(P, A, B, m[0], r[0], e1[0], e2[0], num_keys=P.n - 1)
    ct = O.encrypt_explicit(P, A, B, m[0], r[0], e1[0], e2[0])
    with pytest.raises(O.PvwError):
        O.decrypt_party_shares(P, [], sk[0], 0)
    with pytest.raises(O.PvwError):
        O.decrypt_party_shares(P, [ct], sk[0], 0)            # needs exactly n ciphertexts
    with pytest.raises(O.PvwError):
        O.decrypt_party_shares(P, [ct] * P.n, sk[0], P.n)
    # correctness condition false => encrypt refuses (encryption.rs:124-128)
    Pbad = O.Params(3, 4, 8, [0xFFFFEE001], 0.5, 10 ** 9, 10 ** 9)
    assert not Pbad.verify_correctness_condition()
    # tests/crypto.rs:209-234: bounds (1000, 2000), variance 3 still satisfy it for the 3-moduli set
    assert O.Params(3, 4, 8, O.TEST_MODULI, 3.0, 1000, 2000).verify_correctness_condition()


def test_m_as_i64_wrap_and_u64_rules():
    # encryption.rs:195 (`as i64`) and decryption.rs:226-247
    P = make("T16")
    assert P.encode_scalar(-1) == P.ntt_forward(P.bigints_to_poly([-(P.delta ** i) for i in range(P.l)]))
    assert O._to_u64_rule(P, -1000) == 0 and O._to_u64_rule(P, -1001) == (P.Q - 1001 if P.Q - 1001 < 2 ** 64 else 0)
    assert O._to_u64_rule(P, 2 ** 64) == 0 and O._to_u64_rule(P, 2 ** 64 - 1) == 2 ** 64 - 1


@pytest.mark.parametrize("name", list(CONFIGS))
def test_c_oracle_equals_python_oracle(name):
    P = make(name)
    co = CO.COracle(P)
    D = 2
    A, sk, ke, B, m, r, e1, e2 = _system(P, D, "u63")
    An = CO.synth_crs_np(P)
    assert (np.array(A, dtype=np.uint64) == An).all()
    skn = CO.synth_small_np(P, O.TAG_SK, P.n, P.k, "cbd")
    ken = CO.synth_small_np(P, O.TAG_KE, P.n, P.k, "uniform", P.error_bound_1)
    assert (np.array(sk) == skn).all() and (np.array(ke) == ken).all()
    assert (CO.synth_messages_np(P, D, "u63") == np.array(m, dtype=np.uint64)).all()
    Bn = co.keygen(An, skn, ken)
    assert (np.array(B, dtype=np.uint64) == Bn).all()
    assert (co.ntt_small(np.array(r)) == np.array([[P.ntt_forward(P.from_coefficients(c)) for c in rd] for rd in r],
                                                  dtype=np.uint64)).all()
    cts = [O.encrypt_explicit(P, A, B, m[d], r[d], e1[d], e2[d]) for d in range(D)]
    c1n, c2n = co.encrypt(An, Bn, np.array(m, dtype=np.uint64), r, e1, e2)
    assert (np.array([c[0] for c in cts], dtype=np.uint64) == c1n).all()
    assert (np.array([c[1] for c in cts], dtype=np.uint64) == c2n).all()
    out = co.decrypt(skn, c1n, c2n)
    for p in range(P.n):
        for d in range(D):
            assert int(out[p, d]) == O.decrypt_party_value(P, cts[d][0], cts[d][1], sk[p], p) == m[d][p]
    # failing decodes (uniform garbage) must agree too: literal restatement == scalar form == C
    rng = np.random.default_rng(7)
    zr = np.stack([[rng.integers(0, q, size=P.l, dtype=np.uint64) for q in P.moduli] for _ in range(12)])
    dn = co.decode(zr)
    for i in range(len(zr)):
        zl = [[int(x) for x in row] for row in zr[i]]
        a = O.decode_scalar_pvw_rns(P, zl)
        b = O.decode_scalar_fast(P, P.lift(P.ntt_backward(zl)))
        assert a == b == int(dn[i])
    # lift
    pw = co.ntt_poly(c2n[0, :2], inverse=True)
    lw = co.lift(pw)
    for i in range(2):
        ref = P.lift([[int(x) for x in row] for row in pw[i]])
        got = [sum(int(w) << (64 * t) for t, w in enumerate(lw[i, c])) for c in range(P.l)]
        assert got == ref
