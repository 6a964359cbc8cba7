"""Shared parameter sets and seeded synthetic systems for the parity tests."""
import numpy as np

import c_oracle as CO
import pvw_oracle as O

P128_MODULI = O.largest_ntt_primes(17)           # 17 x 62-bit, Q 1054 bit
P256_MODULI = O.largest_ntt_primes(34)           # 34 x 62-bit, Q 2108 bit (also = 1 mod 32)

CONFIGS = {
    # name: Params kwargs                                                            reference origin
    "EX": dict(n=7, k=32, l=8, moduli=O.EX_MODULI, error_bound_1=50, error_bound_2=50),          # examples/pvw.rs:28-32
    "T16": dict(n=10, k=4, l=16, moduli=O.TEST_MODULI, error_bound_1=50, error_bound_2=50),      # tests/crypto.rs:236-305
    "VDs": dict(n=6, k=40, l=8, moduli=O.VD_MODULI, secret_variance=10.0, error_bound_1=1, error_bound_2=1172385),  # examples/pvw_valid_dec.rs:40-52, k cut
    "L32": dict(n=5, k=3, l=32, moduli=O.largest_ntt_primes(5), secret_variance=1.0),
    "RAG": dict(n=13, k=5, l=8, moduli=O.TEST_MODULI, error_bound_1=50, error_bound_2=50),       # ragged: nothing divides a tile
    "P128s": dict(n=24, k=256, l=8, moduli=P128_MODULI),                                          # 128-bit set, few parties
    "P256s": dict(n=9, k=512, l=16, moduli=P256_MODULI),                                          # 256-bit set, few parties
}


def params(name):
    return O.Params(**CONFIGS[name])


class System:
    """A, sk, B (genuine keys), m, r, e1, e2 from the seeded streams of SURVEY A.7 -- numpy, reference host layout."""

    def __init__(self, P, D, msg_mode="u63", seed=None):
        self.P, self.D = P, D
        self.co = CO.COracle(P)
        self.A = CO.synth_crs_np(P, seed)
        self.sk = CO.synth_small_np(P, O.TAG_SK, P.n, P.k, "cbd", seed=seed)
        self.ke = CO.synth_small_np(P, O.TAG_KE, P.n, P.k, "uniform", P.error_bound_1, seed=seed)
        self.B = self.co.keygen(self.A, self.sk, self.ke)
        self.m = CO.synth_messages_np(P, D, msg_mode, seed=seed)
        self.r = CO.synth_small_np(P, O.TAG_R, D, P.k, "cbd", seed=seed)
        self.e1 = CO.synth_small_np(P, O.TAG_E1, D, P.k, "uniform", P.error_bound_1, seed=seed)
        self.e2 = CO.synth_small_np(P, O.TAG_E2, D, P.n, "uniform", P.error_bound_2, seed=seed)

    def encrypt(self):
        return self.co.encrypt(self.A, self.B, self.m, self.r, self.e1, self.e2)


def engine_kwargs(P, **over):
    kw = dict(n=P.n, k=P.k, l=P.l, moduli=list(P.moduli), psi=list(P.psi), secret_variance=P.secret_variance,
              error_bound_1=P.error_bound_1, error_bound_2=P.error_bound_2)
    kw.update(over)
    return kw
