"""CPU: the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py) against both oracles."""
import os

import numpy as np
import pytest

import c_oracle as CO
import pvw_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["EX", "T16", "RAG"]


def load(name):
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    n, k, l = (int(x) for x in g["nkl"])
    P = O.Params(n, k, l, [int(q) for q in g["moduli"]], float(g["variance"][0]), int(g["bounds"][0]), int(g["bounds"][1]),
                 psi=[int(p) for p in g["psi"]])
    return g, P


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_reproduces_golden(name):
    g, P = load(name)
    assert P.delta == sum(int(w) << (64 * i) for i, w in enumerate(g["delta_words"]))
    assert list(g["psi"]) == [O.fhe_math_default_psi(q, P.l) for q in P.moduli]
    co = CO.COracle(P)
    assert (co.keygen(g["A"], g["sk"], g["ke"]) == g["B"]).all()
    c1, c2 = co.encrypt(g["A"], g["B"], g["m"], g["r"], g["e1"], g["e2"])
    assert (c1 == g["c1"]).all() and (c2 == g["c2"]).all()
    dec, z = co.decrypt(g["sk"], c1, c2, want_zhat=True)
    assert (dec == g["dec"]).all() and (z == g["zhat"]).all()
    assert (co.decode(g["garbage"]) == g["garbage_dec"]).all()


def test_python_oracle_reproduces_golden_ex():
    g, P = load("EX")
    A = g["A"].tolist()
    B = g["B"].tolist()
    c1, c2 = O.encrypt_explicit(P, A, B, [int(x) for x in g["m"][1]], g["r"][1].tolist(), g["e1"][1].tolist(), g["e2"][1].tolist())
    assert (np.array(c1, dtype=np.uint64) == g["c1"][1]).all() and (np.array(c2, dtype=np.uint64) == g["c2"][1]).all()
    for p in range(P.n):
        assert O.decrypt_party_value(P, c1, c2, g["sk"][p].tolist(), p) == int(g["dec"][p, 1]) == int(g["m"][1, p])
    for i in range(4):
        assert O.decode_scalar_pvw_rns(P, g["garbage"][i].tolist()) == int(g["garbage_dec"][i])
