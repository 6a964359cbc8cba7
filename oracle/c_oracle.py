"""ctypes front end of oracle/libpvw_oracle.so (the C restatement) -- TEST INFRASTRUCTURE ONLY.

See the header of pvw_oracle.py / pvw_oracle.c: parity with the real crate is UNPINNED at the fhe-math
boundary; this is the fast checker for sizes the exact-Python oracle cannot reach and the timed CPU baseline.
All arrays use the reference's host layout: polynomial = u64[L][ell] row-major.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpvw_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pvw_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libpvw_oracle.so"])
    return _LIB


class _CParams(C.Structure):
    _fields_ = [("n", C.c_uint32), ("k", C.c_uint32), ("ell", C.c_uint32), ("L", C.c_uint32), ("nw", C.c_uint32),
                ("moduli", C.c_void_p), ("psi", C.c_void_p), ("Q", C.c_void_p), ("delta", C.c_void_p),
                ("delta_pow", C.c_void_p), ("qhat", C.c_void_p), ("qhat_inv", C.c_void_p), ("gadget_rns", C.c_void_p)]


def _words(x: int, nw: int) -> np.ndarray:
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(nw)], dtype=np.uint64)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class COracle:
    """Bound to one oracle `Params` (pvw_oracle.Params)."""

    def __init__(self, P):
        self.lib = C.CDLL(build())
        self.P = P
        L, l = P.L, P.l
        if L > 64 or l > 256:                     # PVWO_MAX_L / PVWO_MAX_ELL of pvw_oracle.c (fixed-size tables)
            raise ValueError(f"the C oracle is built for at most 64 moduli and ring degree 256 (got L={L}, l={l})")
        self.nw = nw = (P.Q.bit_length() + 63) // 64 + 1
        self._keep = dict(
            moduli=_u64(P.moduli), psi=_u64(P.psi), Q=_words(P.Q, nw), delta=_words(P.delta, nw),
            delta_pow=_words(P.delta_power_l_minus_1, nw),
            qhat=np.stack([_words(P.Q // q, nw) for q in P.moduli]),
            qhat_inv=_u64([pow(P.Q // q, -1, q) for q in P.moduli]),
            gadget_rns=_u64([[pow(P.delta, t, q) for t in range(l)] for q in P.moduli]))
        self.cp = _CParams(P.n, P.k, l, L, nw, *[_p(self._keep[f]) for f in
                                                 ("moduli", "psi", "Q", "delta", "delta_pow", "qhat", "qhat_inv", "gadget_rns")])
        self.lib.pvwo_num_threads.restype = C.c_int
        self.poly = (L, l)

    @property
    def threads(self) -> int:
        return int(self.lib.pvwo_num_threads())

    def set_threads(self, n: int):
        """OpenMP threads of the following calls (the reference: rayon's global pool = all host cores)"""
        self.lib.pvwo_set_num_threads.argtypes = [C.c_int]
        self.lib.pvwo_set_num_threads(int(n))

    def ntt_small(self, coef) -> np.ndarray:
        coef = _i64(coef)
        lead = coef.shape[:-1]
        out = np.empty(lead + self.poly, dtype=np.uint64)
        self.lib.pvwo_ntt_small(C.byref(self.cp), C.c_uint64(int(np.prod(lead, dtype=np.int64))), _p(coef), _p(out))
        return out

    def ntt_poly(self, polys, inverse=False) -> np.ndarray:
        polys = _u64(polys).copy()
        cnt = polys.size // (self.poly[0] * self.poly[1])
        self.lib.pvwo_ntt_poly(C.byref(self.cp), C.c_uint64(cnt), _p(polys), C.c_int(1 if inverse else 0))
        return polys

    def encode_scalar(self, m: int) -> np.ndarray:
        out = np.empty(self.poly, dtype=np.uint64)
        self.lib.pvwo_encode_scalar(C.byref(self.cp), C.c_uint64(m), _p(out))
        return out

    def keygen(self, A, sk, e) -> np.ndarray:
        A, sk, e = _u64(A), _i64(sk), _i64(e)
        nparties = sk.shape[0]
        out = np.empty((nparties, self.P.k) + self.poly, dtype=np.uint64)
        self.lib.pvwo_keygen(C.byref(self.cp), C.c_uint64(nparties), _p(A), _p(sk), _p(e), _p(out))
        return out

    def encrypt(self, A, B, m, r, e1, e2, want_c1=True, want_c2=True):
        """m [D][nrows]; r,e1 [D][k][l]; e2 [D][nrows][l]; B [nrows][k][L][l] -> (c1 [D][k][L][l], c2 [D][nrows][L][l])"""
        A, B, m, r, e1, e2 = _u64(A), _u64(B), _u64(m), _i64(r), _i64(e1), _i64(e2)
        D, nrows = m.shape
        c1 = np.empty((D, self.P.k) + self.poly, dtype=np.uint64) if want_c1 else None
        c2 = np.empty((D, nrows) + self.poly, dtype=np.uint64) if want_c2 else None
        self.lib.pvwo_encrypt(C.byref(self.cp), C.c_uint64(D), C.c_uint64(nrows), _p(A), _p(B), _p(m), _p(r), _p(e1), _p(e2),
                              _p(c1) if want_c1 else None, _p(c2) if want_c2 else None)
        return c1, c2

    def decrypt(self, sk, c1, c2, want_zhat=False):
        """sk [P][k][l]; c1 [D][k][L][l]; c2 [D][P][L][l] -> out [P][D] (and zhat [P][D][L][l])"""
        sk, c1, c2 = _i64(sk), _u64(c1), _u64(c2)
        Pn, D = sk.shape[0], c1.shape[0]
        assert c2.shape[:2] == (D, Pn)
        out = np.empty((Pn, D), dtype=np.uint64)
        z = np.empty((Pn, D) + self.poly, dtype=np.uint64) if want_zhat else None
        self.lib.pvwo_decrypt(C.byref(self.cp), C.c_uint64(Pn), C.c_uint64(D), _p(sk), _p(c1), _p(c2), _p(out),
                              _p(z) if want_zhat else None)
        return (out, z) if want_zhat else out

    def decode(self, zhat) -> np.ndarray:
        zhat = _u64(zhat)
        cnt = zhat.size // (self.poly[0] * self.poly[1])
        out = np.empty(cnt, dtype=np.uint64)
        self.lib.pvwo_decode(C.byref(self.cp), C.c_uint64(cnt), _p(zhat), _p(out))
        return out

    def lift(self, polys_power) -> np.ndarray:
        polys_power = _u64(polys_power)
        cnt = polys_power.size // (self.poly[0] * self.poly[1])
        out = np.empty((cnt, self.P.l, self.nw), dtype=np.uint64)
        self.lib.pvwo_lift(C.byref(self.cp), C.c_uint64(cnt), _p(polys_power), _p(out))
        return out


# ------------------------------------------------------------------------------------------------
# vectorised synthetic streams (same values as pvw_oracle.stream_u64 & co)
# ------------------------------------------------------------------------------------------------
def _splitmix64_np(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def stream_np(seed: int, tag: int, start: int, count: int) -> np.ndarray:
    idx = np.arange(start, start + count, dtype=np.uint64)
    base = np.uint64((seed ^ (tag << 56)) & 0xFFFFFFFFFFFFFFFF)
    return _splitmix64_np(idx ^ base)


def _mulhi_np(u: np.ndarray, v) -> np.ndarray:
    """high 64 bits of u * v (v scalar or array), u64 numpy"""
    v = np.asarray(v, dtype=np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    s32 = np.uint64(32)
    with np.errstate(over="ignore"):
        u0, u1 = u & m32, u >> s32
        v0, v1 = v & m32, v >> s32
        t = u0 * v0
        w1 = u1 * v0 + (t >> s32)
        w2 = u0 * v1 + (w1 & m32)
        return u1 * v1 + (w1 >> s32) + (w2 >> s32)


def synth_crs_np(P, seed=None) -> np.ndarray:
    import pvw_oracle as O
    seed = O.DEFAULT_SEED if seed is None else seed
    k, L, l = P.k, P.L, P.l
    u = stream_np(seed, O.TAG_A, 0, k * k * L * l).reshape(k, k, L, l)
    q = _u64(P.moduli).reshape(1, 1, L, 1)
    return _mulhi_np(u, np.broadcast_to(q, u.shape))


def synth_small_np(P, tag, rows, cols, kind, bound=None, seed=None, row0=0) -> np.ndarray:
    import pvw_oracle as O
    seed = O.DEFAULT_SEED if seed is None else seed
    l = P.l
    u = stream_np(seed, tag, row0 * cols * l, rows * cols * l)
    if kind == "cbd":
        v = P.secret_variance
        if abs(v - 0.5) < 1.2e-7:
            out = (u & np.uint64(1)).astype(np.int64) - ((u >> np.uint64(1)) & np.uint64(1)).astype(np.int64)
        else:
            vi = int(v)
            mask = np.uint64((1 << (2 * vi)) - 1)
            pc = lambda a: np.array([bin(int(x)).count("1") for x in a.ravel()], dtype=np.int64).reshape(a.shape) \
                if not hasattr(np, "bitwise_count") else np.bitwise_count(a).astype(np.int64)
            out = pc(u & mask) - pc((u >> np.uint64(2 * vi)) & mask)
    else:
        out = _mulhi_np(u, np.uint64(2 * bound + 1)).astype(np.int64) - np.int64(bound)
    return out.reshape(rows, cols, l)


def synth_messages_np(P, D, mode="example", seed=None, nrows=None, row0=0) -> np.ndarray:
    import pvw_oracle as O
    seed = O.DEFAULT_SEED if seed is None else seed
    n = P.n
    nrows = n if nrows is None else nrows
    d = np.arange(D, dtype=np.uint64).reshape(D, 1)
    p = np.arange(row0, row0 + nrows, dtype=np.uint64).reshape(1, nrows)
    if mode == "example":
        return d * np.uint64(1000) + p + np.uint64(1)
    idx = (d * np.uint64(n) + p)
    base = np.uint64((seed ^ (O.TAG_M << 56)) & 0xFFFFFFFFFFFFFFFF)
    return _splitmix64_np(idx ^ base) >> np.uint64(1)
