"""Generates tests/golden/*.npz from the oracle (exact-Python oracle cross-checked against the C restatement).

The reference (Rust + un-vendored fhe-math) cannot be built or imported in this image and holds no known-answer
vectors, so these fixtures pin OUR restatement (parity unpinned at the fhe-math boundary, see oracle/pvw_oracle.py);
they exist so that the GPU path, the C oracle and the Python oracle are all held to the same committed bytes.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import pvw_oracle as O  # noqa: E402
from _cases import CONFIGS, System, params  # noqa: E402


def make(name, D, msg_mode):
    P = params(name)
    S = System(P, D, msg_mode)
    c1, c2 = S.encrypt()
    # exact-Python oracle on the same inputs (small configs only)
    A = [[[[int(x) for x in row] for row in poly] for poly in r] for r in S.A]
    B = [[[[int(x) for x in row] for row in poly] for poly in r] for r in S.B]
    for d in range(D):
        p1, p2 = O.encrypt_explicit(P, A, B, [int(x) for x in S.m[d]], S.r[d].tolist(), S.e1[d].tolist(), S.e2[d].tolist())
        assert (np.array(p1, dtype=np.uint64) == c1[d]).all() and (np.array(p2, dtype=np.uint64) == c2[d]).all()
    dec, zhat = S.co.decrypt(S.sk, c1, c2, want_zhat=True)
    for p in range(P.n):
        for d in range(D):
            assert int(dec[p, d]) == O.decrypt_party_value(P, p1 if d == D - 1 else O.encrypt_explicit(
                P, A, B, [int(x) for x in S.m[d]], S.r[d].tolist(), S.e1[d].tolist(), S.e2[d].tolist())[0],
                [[[int(x) for x in row] for row in poly] for poly in c2[d]], S.sk[p].tolist(), p)
    rng = np.random.default_rng(20261018)
    garbage = np.stack([[rng.integers(0, q, size=P.l, dtype=np.uint64) for q in P.moduli] for _ in range(16)])
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), moduli=np.array(P.moduli, dtype=np.uint64), psi=np.array(P.psi, dtype=np.uint64),
                        nkl=np.array([P.n, P.k, P.l]), bounds=np.array([P.error_bound_1, P.error_bound_2]),
                        variance=np.array([P.secret_variance]), A=S.A, sk=S.sk, ke=S.ke, B=S.B, m=S.m, r=S.r, e1=S.e1, e2=S.e2,
                        c1=c1, c2=c2, dec=dec, zhat=zhat, garbage=garbage, garbage_dec=S.co.decode(garbage),
                        delta_words=np.array([(P.delta >> (64 * i)) & (2 ** 64 - 1) for i in range(4)], dtype=np.uint64))
    print(name, "ok", os.path.getsize(os.path.join(HERE, f"{name}.npz")), "bytes")


def make_wire():
    """wire-format blobs (oracle/pvw_wire.py; third-party encodings recalled -- parity unpinned) of the EX system"""
    import pvw_wire as W
    P = params("EX")
    S = System(P, 2)
    c1, c2 = S.encrypt()
    u8 = lambda b: np.frombuffer(b, dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "EX_wire.npz"), ct1=u8(W.ciphertext_to_bytes(P, c1[1].tolist(), c2[1].tolist())),
                        params=u8(W.params_to_bytes(P)), pk3=u8(W.public_key_to_bytes(P, S.B[3].tolist())),
                        crs=u8(W.crs_to_bytes(P, S.A.tolist())))
    print("EX_wire ok", os.path.getsize(os.path.join(HERE, "EX_wire.npz")), "bytes")


if __name__ == "__main__":
    if sys.argv[1:] == ["wire"]:
        make_wire()
        sys.exit(0)
    make_wire()
    make("EX", 3, "example")
    make("T16", 4, "u63")
    make("RAG", 3, "u63")
