"""GPU: the one-kernel decode of clean shares (csrc/decode.cu (0)) against the oracle on hand-built noisy polynomials.

The fast path claims results identical to decode_scalar_pvw_rns (src/crypto/decryption.rs:10-58 with helpers :61-247) for EVERY
input: shares whose integers are all small are decoded in place, any share that fails one of its checks goes to the general
three-kernel chain.  The cases below sit on both sides of every check: noise at +-Delta/2 and one off (the centred-remainder tie
and the exact-divisibility test), a first-coefficient error that makes `last` wrap modulo Q (the cmax bound), negative and
oversized plaintexts (the u64 rules of :226-247), values too large for the short CRT lift, and uniform garbage.  Each set is
decoded with the fast path on and off and compared with the C restatement (and with the exact Python restatement on a sample)."""
import numpy as np
import pytest

import c_oracle as CO
import pvw_oracle as O
from _cases import engine_kwargs, params

pytestmark = pytest.mark.gpu

SETS = ["VDs", "P128s", "P256s", "L32", "EX", "T16"]      # EX / T16: no sub-basis exists, the fast path is off by construction


def noisy_poly(P, m_signed, e):
    """NTT form of z with z_i = -(m D^i + e_i) mod Q  (decryption.rs:274: noisy = <s, c1> - c2 = -(m g + noise))"""
    coeffs = [(-(m_signed * P.delta ** i + e[i])) % P.Q for i in range(P.l)]
    return P.ntt_forward(P.bigints_to_poly(coeffs))


def cases(P, rng):
    l, D, Q, M = P.l, P.delta, P.Q, P.delta_power_l_minus_1
    h = D // 2
    out = []
    small = lambda: [int(x) for x in rng.integers(-1000, 1001, size=l)]
    msgs = [0, 1, 12345678901234567, (1 << 63) - 1, (1 << 62) + 7]
    for m in msgs:
        out.append((m, small()))
    out.append((-3, small()))                      # m >= 2^63 encodes a negative scalar (encryption.rs:195): small negative -> 0
    out.append((-1000, [0] * l))
    out.append((-1001, [0] * l))                   # |pt| > 1000: (pt + Q) % Q, to_u64().unwrap_or(0)
    out.append((-(1 << 62), small()))
    for pos in (0, 1, l // 2, l - 2, l - 1):       # one coefficient at / around the rounding boundary
        for v in (h, -h, h + 1, -(h + 1), h - 1, -(h - 1), D, -D, 3 * D + 1):
            e = small()
            e[pos] = v
            out.append((424242, e))
    out.append((7, [h] * l))
    out.append((7, [-h] * l))
    out.append((7, [h + 1] * l))
    # |H| ~ |e_0| D^(l-1) around Q/2: `last` wraps modulo Q on one side of the bound
    c = Q // (2 * M)
    for v in (c - 2, c - 1, c, c + 1, c + 2, -(c - 1), -c, -(c + 1)):
        e = small()
        e[0] = v
        out.append((99, e))
    # values beyond the sub-basis of the short lift (huge noise everywhere), and plain garbage
    for _ in range(6):
        out.append((5, [int(rng.integers(-(1 << 62), 1 << 62)) * (1 << int(rng.integers(0, max(1, D.bit_length() + 40)))) for _ in range(l)]))
    return out


@pytest.mark.parametrize("name", SETS)
def test_fused_decode_matches_the_oracle_on_both_sides_of_every_check(name):
    import pvw_rs_b200 as pvw
    P = params(name)
    rng = np.random.default_rng(17)
    cs = cases(P, rng)
    polys = [noisy_poly(P, m, e) for m, e in cs]
    zhat = np.array(polys, dtype=np.uint64)                                   # [count][L][l]
    garbage = np.stack([rng.integers(0, q, size=(64, P.l), dtype=np.uint64) for q in P.moduli], axis=1)
    zhat = np.concatenate([zhat, garbage])
    # many copies so that every thread group of the kernel sees a mix of clean and handed-over shares
    reps = rng.permutation(np.tile(np.arange(len(zhat)), 7))
    batch = np.ascontiguousarray(zhat[reps])
    co = CO.COracle(P)
    want = co.decode(batch)
    for i in rng.choice(len(cs), size=min(12, len(cs)), replace=False):       # the exact (literal) restatement agrees with the C one
        assert O.decode_scalar_pvw_rns(P, polys[i]) == int(co.decode(zhat[i:i + 1])[0])
    eng = pvw.Engine(**engine_kwargs(P))
    for fused in (1, 0):                                  # fast path + per-share fallback, general chain only
        eng.set_option("decode_fused", fused)
        eng.set_option("profile", 2)
        got = eng.decode_batch(batch)
        prof = eng.profile()
        eng.set_option("profile", 0)
        assert (got == want).all(), f"decode_fused={fused}: {(got != want).sum()} of {len(want)} differ"
        if fused == 0:
            assert prof["decode_fused"][1] == 0
    if name in ("VDs", "P128s", "P256s", "L32"):
        eng.set_option("decode_fused", 1)
        eng.set_option("profile", 2)
        eng.decode_batch(batch[:8])
        assert eng.profile()["decode_fused"][1] == 1                          # the fast path exists for these sets and ran
        eng.set_option("profile", 0)
    # clean shares decode to their messages
    for (m, e), v in zip(cs[:5], co.decode(zhat[:5])):
        assert int(v) == m
