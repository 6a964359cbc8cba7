"""Host-side mirror of the pvw-rs API for the hot path: same names, argument meaning and error behaviour as the
crate (src/lib.rs:31-55), every ring operation forwarded to the CUDA kernels through the C ABI.

Differences a caller sees (all stated in DESIGN.md):
  * `encrypt*` accept an optional `randomness=(r, e1, e2)`; without it r / e1 / e2 are sampled on the host with the
    reference's distributions (src/sampling/uniform.rs) -- the reference draws them from thread_rng() and is not
    reproducible (src/crypto/encryption.rs:138,164,180).
  * polynomials are numpy u64 arrays of shape (L, l) in the reference's row-major layout instead of fhe-math `Poly`.
  * the public key matrix B, the CRS A and ciphertexts live on the device; `.matrix`, `.c1`, `.c2` download on access.
"""
from __future__ import annotations

import threading
import weakref
from typing import List, Optional, Sequence

import numpy as np

from .engine import Engine
from .errors import PvwError

_U64 = (1 << 64) - 1


# --------------------------------------------------------------------------------------------------------------
# sampling (host) -- src/sampling/uniform.rs
# --------------------------------------------------------------------------------------------------------------
class _OsRng:
    """Uniform integers drawn from the operating system's CSPRNG (os.urandom = getrandom(2)).  The reference samples keys,
    randomness and errors from `thread_rng()`, a CryptoRng (ChaCha12 seeded from the OS); numpy's generators (PCG64,
    MT19937, ...) are not cryptographic and must never produce secret material."""

    def integers(self, low, high=None, size=None, dtype=np.int64):
        import os
        if high is None:
            low, high = 0, low
        low, high = int(low), int(high)
        span = high - low
        if span <= 0:
            raise ValueError("empty range")
        shape = () if size is None else ((size,) if np.isscalar(size) else tuple(size))
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if span & (span - 1) == 0 and span <= 256:                  # power of two <= 256: one byte each, no rejection
            x = np.frombuffer(os.urandom(count), dtype=np.uint8).astype(np.uint64) & np.uint64(span - 1)
        else:
            wdt, bits = (np.uint32, 32) if span <= (1 << 16) else (np.uint64, 64)
            limit = ((1 << bits) // span) * span                      # rejection sampling: accept draws below the largest multiple of span
            x = np.empty(count, dtype=np.uint64)
            todo = np.arange(count)
            while len(todo):
                draw = np.frombuffer(os.urandom(len(todo) * (bits // 8)), dtype=wdt).astype(np.uint64)
                ok = draw < np.uint64(limit) if limit < (1 << 64) else np.ones(len(draw), dtype=bool)
                x[todo[ok]] = draw[ok] % np.uint64(span)
                todo = todo[~ok]
        if low >= 0 and high <= (1 << 64) and np.dtype(dtype) == np.uint64:
            out = x + np.uint64(low)
        else:
            out = (x.astype(np.int64) + np.int64(low)).astype(dtype)
        return out.reshape(shape) if shape else out[0]


def _rng(rng):
    """None -> the OS CSPRNG (production).  A numpy Generator or an integer seed gives a reproducible NON-cryptographic stream:
    for tests and fixtures only (the reference has no such option; its samplers take `RngCore + CryptoRng`)."""
    if rng is None:
        return _OsRng()
    return rng if isinstance(rng, (np.random.Generator, _OsRng)) else np.random.default_rng(rng)


def sample_uniform_coefficients(bound: int, num_coeffs: int, rng=None) -> np.ndarray:
    """uniform integers in [-bound, bound] (uniform.rs:5-22)"""
    return _rng(rng).integers(-int(bound), int(bound) + 1, size=num_coeffs, dtype=np.int64)


def sample_vec_cbd(vector_size: int, variance: float, rng=None) -> np.ndarray:
    """centred binomial distribution (uniform.rs:27-70): variance 0.5 -> {-1,0,1}; integer variance v -> 2v + 2v bits"""
    if variance <= 0:
        raise PvwError("SamplingError", "The variance should be positive")
    g = _rng(rng)
    if abs(variance - 0.5) < 1.2e-7:
        bits = g.integers(0, 4, size=vector_size, dtype=np.int64)
        return (bits & 1) - ((bits >> 1) & 1)
    v = int(variance)
    if v < 1 or v > 16 or v != variance:
        raise PvwError("SamplingError", "The variance should be an integer between 1 and 16")
    a = g.integers(0, 2, size=(vector_size, 2 * v), dtype=np.int64).sum(axis=1)
    b = g.integers(0, 2, size=(vector_size, 2 * v), dtype=np.int64).sum(axis=1)
    return a - b


# --------------------------------------------------------------------------------------------------------------
# parameters -- src/params/parameters.rs
# --------------------------------------------------------------------------------------------------------------
class PvwParameters:
    """PvwParameters (parameters.rs:19-40).  Validation and the derived integers (Delta, Delta^(l-1), Q) come from the
    library's context constructor, which restates PvwParametersBuilder::build (parameters.rs:117-195)."""

    def __init__(self, n: int, k: int, l: int, moduli: Sequence[int], secret_variance: float = 0.5,
                 error_bound_1: int = 100, error_bound_2: int = 200, psi: Optional[Sequence[int]] = None, device: int = 0):
        if int(error_bound_1) <= 0:
            raise PvwError("InvalidParameters", "error_bound_1 must be positive")
        if int(error_bound_2) <= 0:
            raise PvwError("InvalidParameters", "error_bound_2 must be positive")
        self._engine_args = dict(n=n, k=k, l=l, moduli=list(moduli), psi=list(psi) if psi is not None else None,
                                 secret_variance=secret_variance, error_bound_1=error_bound_1, error_bound_2=error_bound_2,
                                 device=device)
        eng = Engine(**self._engine_args)          # validates exactly like the builder + fhe-math Context::new
        self._probe = eng
        self.n, self.k, self.l = int(n), int(k), int(l)
        self.t = (self.n - 1) // 2                  # parameters.rs:169
        self.secret_variance = float(secret_variance)
        self.error_bound_1, self.error_bound_2 = int(error_bound_1), int(error_bound_2)
        self._moduli = [int(q) for q in moduli]
        self.psi = eng.psi
        self.delta = eng.delta
        self.delta_power_l_minus_1 = eng.delta_power_l_minus_1
        self._q_total = eng.q_total

    @staticmethod
    def builder() -> "PvwParametersBuilder":
        return PvwParametersBuilder()

    @property
    def L(self) -> int:
        return len(self._moduli)

    def moduli(self) -> List[int]:
        return list(self._moduli)

    def q_total(self) -> int:
        return self._q_total

    def new_engine(self, row0: int = 0, nrows: int = 0, device: Optional[int] = None) -> Engine:
        args = dict(self._engine_args)
        args["psi"] = self.psi
        if device is not None:
            args["device"] = device
        return Engine(row0=row0, nrows=nrows, **args)

    def verify_correctness_condition(self) -> bool:
        return self._probe.verify_correctness_condition()

    def gadget_vector(self) -> List[int]:
        return [self.delta ** i for i in range(self.l)]                               # parameters.rs:311-325

    def bigints_to_poly(self, bigints: Sequence[int]) -> np.ndarray:
        """parameters.rs:420-474 -- PowerBasis polynomial, residues ((x % q) + q) % q (host-side integer reduction)"""
        if len(bigints) != self.l:
            raise PvwError("InvalidParameters", f"Expected {self.l} coefficients, got {len(bigints)}")
        return np.array([[int(x) % q for x in bigints] for q in self._moduli], dtype=np.uint64)

    def ntt_forward_small(self, coeffs) -> np.ndarray:
        """Poly::from_coefficients(&[i64]) + change_representation(Ntt) on the device"""
        return self._probe.ntt_forward_small(coeffs)

    def encode_scalar(self, scalar: int) -> np.ndarray:
        """parameters.rs:346-367, Ntt form: NTT(scalar * [1, D, ..., D^(l-1)]) == scalar * NTT(gadget)"""
        if not -(1 << 63) <= int(scalar) < (1 << 63):
            raise PvwError("InvalidParameters", "scalar does not fit i64")
        return self._probe.encode_scalars([int(scalar) & _U64])[0]

    @staticmethod
    def suggest_error_bounds(n: int, k: int, l: int, moduli: Sequence[int], secret_variance: float):
        """parameters.rs:554-603"""
        import math
        tmp = PvwParameters(n, k, l, moduli, secret_variance, 1, 1)
        try:
            dp = float(tmp.delta_power_l_minus_1)
        except OverflowError:
            dp = math.inf
        c1 = 2.0 * k * l + 14.0 * math.sqrt(float(n) * k * l)
        c2 = math.sqrt(float(n) * l) * (1.0 + math.sqrt(float(n)))
        for b1 in (50, 100, 200, 500, 1000, 2000):
            for b2 in (50, 100, 200, 500, 1000, 2000):
                if dp > b1 * c1 + b2 * c2:
                    return b1, b2
        raise PvwError("InvalidParameters", "Cannot find suitable error bounds")


class PvwParametersBuilder:
    """parameters.rs:44-201; defaults variance 0.5, bounds 100 / 200 (:166-168)"""

    def __init__(self):
        self._n = self._k = self._l = None
        self._moduli = None
        self._variance, self._b1, self._b2, self._psi, self._device = 0.5, 100, 200, None, 0

    def set_parties(self, n): self._n = n; return self
    def set_dimension(self, k): self._k = k; return self
    def set_l(self, l): self._l = l; return self
    def set_moduli(self, moduli): self._moduli = list(moduli); return self
    def set_secret_variance(self, v): self._variance = v; return self
    def set_error_bound_1(self, b): self._b1 = b; return self
    def set_error_bound_2(self, b): self._b2 = b; return self
    def set_error_bounds(self, b1, b2): self._b1, self._b2 = b1, b2; return self
    def set_error_bounds_u32(self, b1, b2): return self.set_error_bounds(b1, b2)
    def set_psi(self, psi): self._psi = list(psi); return self
    def set_device(self, d): self._device = d; return self

    def build(self) -> PvwParameters:
        for name, v in (("n", self._n), ("k", self._k), ("l", self._l), ("moduli", self._moduli)):
            if v is None:
                raise PvwError("InvalidParameters", f"{name} not set")
        return PvwParameters(self._n, self._k, self._l, self._moduli, self._variance, self._b1, self._b2, self._psi, self._device)

    build_arc = build


# --------------------------------------------------------------------------------------------------------------
# CRS -- src/params/crs.rs
# --------------------------------------------------------------------------------------------------------------
def expand_crs_seed(params: "PvwParameters", seed: bytes) -> np.ndarray:
    from . import _ffi
    seed = bytes(seed)
    if len(seed) != 32:
        raise PvwError("InvalidParameters", "the master seed must be 32 bytes")
    mods = np.array(params.moduli(), dtype=np.uint64)
    out = np.empty((params.k, params.k, params.L, params.l), dtype=np.uint64)
    rc = _ffi.load().pvw_crs_expand_seed(params.k, params.l, params.L, mods.ctypes.data, seed, out.ctypes.data)
    if rc != 0:
        raise PvwError(_ffi.STATUS_NAMES.get(rc, "InternalError"), "CRS seed expansion failed")
    return out


class PvwCrs:
    """PvwCrs{matrix: k x k polynomials in NTT form, params} (crs.rs:12-17); matrix is a host array [k][k][L][l]."""

    def __init__(self, params: PvwParameters, matrix: np.ndarray):
        shape = (params.k, params.k, params.L, params.l)
        matrix = np.ascontiguousarray(matrix, dtype=np.uint64)
        if matrix.shape != shape:
            raise PvwError("DimensionMismatch", f"expected {shape}, got {matrix.shape}")
        self.params, self.matrix = params, matrix
        self._engine: Optional[Engine] = None

    @classmethod
    def new(cls, params: PvwParameters, rng=None) -> "PvwCrs":
        """crs.rs:24-39: uniform residues labelled NTT (host RNG; stream differs from rand's, distribution is the same)"""
        g = _rng(rng)
        m = np.empty((params.k, params.k, params.L, params.l), dtype=np.uint64)
        for j, q in enumerate(params.moduli()):
            m[:, :, j, :] = g.integers(0, q, size=(params.k, params.k, params.l), dtype=np.uint64)
        return cls(params, m)

    @classmethod
    def new_deterministic(cls, params: PvwParameters, seed: bytes) -> "PvwCrs":
        """crs.rs:45-67: same 32-byte seed => identical CRS (host-side expansion in the library, no device needed)"""
        return cls(params, expand_crs_seed(params, seed))

    @classmethod
    def new_from_tag(cls, params: PvwParameters, tag: str) -> "PvwCrs":
        """crs.rs:74-90"""
        import ctypes
        from . import _ffi
        seed = (ctypes.c_uint8 * 32)()
        rc = _ffi.load().pvw_crs_tag_to_seed(tag.encode(), seed)
        if rc != 0:
            raise PvwError(_ffi.STATUS_NAMES.get(rc, "InternalError"), "tag")
        return cls.new_deterministic(params, bytes(seed))

    def dimensions(self):
        return self.matrix.shape[0], self.matrix.shape[1]

    def get(self, i, j):
        return self.matrix[i, j] if 0 <= i < self.params.k and 0 <= j < self.params.k else None

    def __len__(self):
        return self.params.k * self.params.k

    def validate(self):
        if self.dimensions() != (self.params.k, self.params.k):
            raise PvwError("InvalidParameters", "CRS dimensions mismatch")

    def _eng(self) -> Engine:
        if self._engine is None:
            self._engine = self.params.new_engine(0, 1)
            self._engine.crs_upload(self.matrix)
        return self._engine

    def multiply_by_randomness(self, randomness) -> np.ndarray:
        """crs.rs:177-205: out[i] = sum_j A[i][j] * r[j]; randomness = k NTT-form polynomials"""
        r = np.asarray(randomness, dtype=np.uint64)
        if r.shape[0] != self.params.k:
            raise PvwError("DimensionMismatch", f"expected {self.params.k}, actual {r.shape[0]}")
        return self._eng().crs_multiply_by_randomness(r[None])[0]

    def multiply_by_secret_key(self, secret_key: "SecretKey") -> np.ndarray:
        """crs.rs:138-171: result[i] = sum_j NTT(s_j) * A[j][i]  (= the error-free public key)"""
        eng = self._eng()
        eng.keygen_batch(0, secret_key.secret_coeffs[None], np.zeros((1, self.params.k, self.params.l), dtype=np.int64))
        return eng.pk_download_rows(0, 1)[0]


# --------------------------------------------------------------------------------------------------------------
# keys -- src/keys/secret_key.rs, src/keys/public_key.rs
# --------------------------------------------------------------------------------------------------------------
class SecretKey:
    """SecretKey{params, secret_coeffs: k x l i64} (secret_key.rs:14-18)"""

    def __init__(self, params: PvwParameters, secret_coeffs):
        c = np.ascontiguousarray(secret_coeffs, dtype=np.int64)
        if c.shape != (params.k, params.l):                                            # from_coefficients, secret_key.rs:258-269
            raise PvwError("InvalidParameters", f"Expected {params.k}x{params.l} coefficients, got {c.shape}")
        self.params, self.secret_coeffs = params, c

    @classmethod
    def random(cls, params: PvwParameters, rng=None) -> "SecretKey":
        """secret_key.rs:45-63: k vectors of l CBD coefficients"""
        return cls(params, sample_vec_cbd(params.k * params.l, params.secret_variance, rng).reshape(params.k, params.l))

    from_coefficients = classmethod(lambda cls, params, coeffs: cls(params, coeffs))

    def coefficients(self):
        return self.secret_coeffs

    def __len__(self):
        return self.params.k

    def get_polynomial(self, index: int) -> np.ndarray:
        """secret_key.rs:98-112: from_coefficients + forward NTT"""
        if not 0 <= index < self.params.k:
            raise PvwError("InvalidParameters", f"Index {index} out of bounds")
        return self.params.ntt_forward_small(self.secret_coeffs[index])

    def to_polynomials(self) -> np.ndarray:
        return self.params.ntt_forward_small(self.secret_coeffs)


class PublicKey:
    """PublicKey{key_polynomials: k polys (NTT), params} (public_key.rs:30-40)"""

    def __init__(self, params: PvwParameters, key_polynomials: np.ndarray):
        self.params, self.key_polynomials = params, np.ascontiguousarray(key_polynomials, dtype=np.uint64)

    @classmethod
    def generate(cls, secret_key: SecretKey, crs: PvwCrs, rng=None, errors=None) -> "PublicKey":
        """public_key.rs:111-147: b = s*A + e with e uniform in [-error_bound_1, error_bound_1]"""
        P = secret_key.params
        e = sample_uniform_coefficients(P.error_bound_1, P.k * P.l, rng).reshape(P.k, P.l) if errors is None else np.asarray(errors, np.int64)
        eng = crs._eng()
        eng.keygen_batch(0, secret_key.secret_coeffs[None], e[None])
        return cls(P, eng.pk_download_rows(0, 1)[0])

    def dimension(self):
        return len(self.key_polynomials)

    def validate(self):
        if len(self.key_polynomials) != self.params.k:
            raise PvwError("InvalidParameters", "Public key dimension mismatch")


class Party:
    """Party{index, secret_key} (public_key.rs:17-27, 62-79)"""

    def __init__(self, index: int, params: PvwParameters, rng=None, secret_key: Optional[SecretKey] = None):
        if index >= params.n:
            raise PvwError("InvalidParameters", f"Party index {index} exceeds maximum {params.n - 1}")
        self._index = index
        self._sk = secret_key if secret_key is not None else SecretKey.random(params, rng)

    new = classmethod(lambda cls, index, params, rng=None: cls(index, params, rng))

    def index(self):
        return self._index

    def secret_key(self) -> SecretKey:
        return self._sk


class GlobalPublicKey:
    """GlobalPublicKey{matrix n x k (NTT), crs, num_keys, params} (public_key.rs:43-54).  The matrix B and the CRS A are
    device resident (one Engine per key); `ct_capacity` ciphertext slots are reserved for encrypt / decrypt."""

    def __init__(self, crs: PvwCrs, ct_capacity: Optional[int] = None):
        self.crs_ = crs
        self.params = crs.params
        self.engine = self.params.new_engine(0, 0)
        self.engine.crs_upload(crs.matrix)
        self._slots = _SlotPool(self.engine, ct_capacity if ct_capacity is not None else max(self.params.n, 4))
        self._lock = threading.RLock()
        self.error_polynomials: List[np.ndarray] = []      # per party: [k][L][l] NTT-form errors, or an empty array (public_key.rs:53)

    new = classmethod(lambda cls, crs: cls(crs))

    @property
    def num_keys(self) -> int:
        return self.engine.num_keys

    @property
    def matrix(self) -> np.ndarray:
        return self.engine.pk_download_rows(0, self.params.n)

    def crs(self) -> PvwCrs:
        return self.crs_

    def dimensions(self):
        return self.params.n, self.params.k

    def num_public_keys(self) -> int:
        return self.num_keys

    def is_full(self) -> bool:
        return self.num_keys >= self.params.n                                           # public_key.rs:349-351

    def add_public_key(self, index: int, public_key: PublicKey):
        """public_key.rs:214-250"""
        if index >= self.params.n:
            raise PvwError("IndexOutOfBounds", f"Party index {index} exceeds maximum {self.params.n - 1}")
        public_key.validate()
        with self._lock:
            self.engine.pk_upload_rows(index, public_key.key_polynomials[None])

    def generate_and_add_party(self, party: Party, rng=None, errors=None):
        """public_key.rs:256-263 -- the row is generated on the device straight into B"""
        P = self.params
        e = sample_uniform_coefficients(P.error_bound_1, P.k * P.l, rng).reshape(P.k, P.l) if errors is None else np.asarray(errors, np.int64)
        with self._lock:
            self.engine.keygen_batch(party.index(), party.secret_key().secret_coeffs[None], e[None])

    def generate_and_add(self, index: int, secret_key: SecretKey, rng=None, errors=None):
        """public_key.rs:268-277"""
        self.generate_and_add_party(Party(index, self.params, secret_key=secret_key), rng, errors)

    def generate_and_add_with_errors(self, index: int, secret_key: SecretKey, rng=None, errors=None):
        """public_key.rs:304-321: like generate_and_add, and the NTT-form error polynomials are kept"""
        P = self.params
        if index >= P.n:
            raise PvwError("IndexOutOfBounds", f"Party index {index} exceeds maximum {P.n - 1}")
        e = sample_uniform_coefficients(P.error_bound_1, P.k * P.l, rng).reshape(P.k, P.l) if errors is None else np.asarray(errors, np.int64)
        with self._lock:
            self.engine.keygen_batch(index, secret_key.secret_coeffs[None], e[None])
            while len(self.error_polynomials) <= index:
                self.error_polynomials.append(np.zeros((0, P.L, P.l), dtype=np.uint64))
            self.error_polynomials[index] = self.engine.ntt_forward_small(e)

    def generate_and_add_party_with_errors(self, party: Party, rng=None, errors=None):
        """public_key.rs:322-329"""
        self.generate_and_add_with_errors(party.index(), party.secret_key(), rng, errors)

    def get_party_errors(self, party_index: int):
        return self.error_polynomials[party_index] if 0 <= party_index < len(self.error_polynomials) else None

    def get_all_errors(self):
        return self.error_polynomials

    def generate_all_party_keys(self, parties: Sequence[Party], rng=None, errors=None):
        """public_key.rs:376-401: every key lands in row party.index() (add_public_key, :214-250), whatever the order of the list;
        one batched device keygen per run of consecutive indices (a single call for the usual 0..len-1 list)"""
        P = self.params
        if len(parties) > P.n:
            raise PvwError("InvalidParameters", f"Too many parties: {len(parties)} > {P.n}")
        for p in parties:
            if not 0 <= p.index() < P.n:
                raise PvwError("IndexOutOfBounds", f"Party index {p.index()} exceeds maximum {P.n - 1}")
        sk = np.stack([p.secret_key().secret_coeffs for p in parties]) if parties else np.zeros((0, P.k, P.l), np.int64)
        e = (sample_uniform_coefficients(P.error_bound_1, len(parties) * P.k * P.l, rng).reshape(len(parties), P.k, P.l)
             if errors is None else np.asarray(errors, np.int64))
        if e.shape != sk.shape:
            raise PvwError("DimensionMismatch", f"errors: expected {sk.shape}, got {e.shape}")
        idx = [p.index() for p in parties]
        with self._lock:
            start = 0
            while start < len(idx):
                stop = start + 1
                while stop < len(idx) and idx[stop] == idx[stop - 1] + 1:
                    stop += 1
                self.engine.keygen_batch(idx[start], sk[start:stop], e[start:stop])
                start = stop

    def generate_all_keys(self, secret_keys: Sequence[SecretKey], rng=None, errors=None):
        """public_key.rs:407-434"""
        self.generate_all_party_keys([Party(i, self.params, secret_key=s) for i, s in enumerate(secret_keys)], rng, errors)

    def get_public_key(self, index: int) -> Optional[PublicKey]:
        if index >= self.num_keys:
            return None
        return PublicKey(self.params, self.engine.pk_download_rows(index, 1)[0])

    def get_polynomial(self, i: int, j: int):
        if i >= self.params.n or j >= self.params.k:
            return None
        return self.engine.pk_download_rows(i, 1)[0, j]

    def validate(self):
        if self.num_keys > self.params.n:
            raise PvwError("InvalidParameters", "Too many public keys")


# --------------------------------------------------------------------------------------------------------------
# ciphertexts -- src/crypto/encryption.rs
# --------------------------------------------------------------------------------------------------------------
class _SlotPool:
    """device ciphertext slots; when none is free the oldest resident ciphertext is spilled to host arrays"""

    def __init__(self, engine: Engine, capacity: int):
        self.engine = engine
        engine.ct_reserve(capacity)
        self.free = list(range(capacity - 1, -1, -1))
        self.resident: "dict[int, weakref.ref]" = {}
        self.order: List[int] = []

    def take(self, owner: "PvwCiphertext", pinned=()) -> int:
        """a slot for `owner`; slots in `pinned` (ciphertexts of the call in progress) are never spilled"""
        if not self.free:
            for s in list(self.order):
                if s in pinned:
                    continue
                ct = self.resident[s]()
                if ct is not None:
                    ct._spill()
                else:
                    self._release(s)
                if self.free:
                    break
        if not self.free:
            raise PvwError("InvalidParameters", "every ciphertext slot is pinned by the call in progress")
        s = self.free.pop()
        self.resident[s] = weakref.ref(owner)
        self.order.append(s)
        return s

    def take_many(self, owners: Sequence["PvwCiphertext"]) -> List[int]:
        """contiguous run of slots for a batched encrypt"""
        n = len(owners)
        if n > self.engine.capacity:
            raise PvwError("InvalidParameters", f"batch of {n} ciphertexts exceeds the reserved capacity {self.engine.capacity}")
        for s in list(self.order):
            if s < n:
                ct = self.resident[s]()
                if ct is not None:
                    ct._spill()
                else:
                    self._release(s)
        for s in range(n):
            self.free.remove(s)
            self.resident[s] = weakref.ref(owners[s])
            self.order.append(s)
        return list(range(n))

    def _release(self, s: int):
        if s in self.resident:
            del self.resident[s]
            self.order.remove(s)
            self.free.append(s)


class PvwCiphertext:
    """PvwCiphertext{c1: k polys, c2: n polys, params} (encryption.rs:15-24), device resident with lazy host copies"""

    def __init__(self, global_pk: GlobalPublicKey, c1: Optional[np.ndarray] = None, c2: Optional[np.ndarray] = None):
        self.params = global_pk.params
        self._pk = global_pk
        self._slot: Optional[int] = None
        self._c1, self._c2 = c1, c2

    def _spill(self):
        if self._slot is not None:
            c1, c2 = self._pk.engine.ct_download(self._slot, self._c1 is None, self._c2 is None)
            self._c1 = c1 if self._c1 is None else self._c1
            self._c2 = c2 if self._c2 is None else self._c2
            self._pk._slots._release(self._slot)
            self._slot = None

    def _resident_slot(self, pinned=()) -> int:
        if self._slot is None:
            self._slot = self._pk._slots.take(self, pinned)
            self._pk.engine.ct_upload(self._slot, self._c1, self._c2)
        return self._slot

    def __del__(self):
        try:
            if self._slot is not None:
                self._pk._slots._release(self._slot)
        except Exception:
            pass

    @property
    def c1(self) -> np.ndarray:
        if self._c1 is None:
            self._c1 = self._pk.engine.ct_download(self._slot, True, False)[0]
        return self._c1

    @property
    def c2(self) -> np.ndarray:
        if self._c2 is None:
            self._c2 = self._pk.engine.ct_download(self._slot, False, True)[1]
        return self._c2

    def c1_components(self):
        return self.c1

    def c2_components(self):
        return self.c2

    def __len__(self):
        return self.params.n

    def get_party_ciphertext(self, party_index: int):
        return self.c2[party_index] if 0 <= party_index < self.params.n else None

    def validate(self):
        """encryption.rs:41-76"""
        if self._c1 is not None and len(self._c1) != self.params.k:
            raise PvwError("InvalidParameters", f"c1 length {len(self._c1)} != k {self.params.k}")
        if self._c2 is not None and len(self._c2) != self.params.n:
            raise PvwError("InvalidParameters", f"c2 length {len(self._c2)} != n {self.params.n}")


def _sample_randomness(P: PvwParameters, D: int, rng=None):
    g = _rng(rng)
    r = sample_vec_cbd(D * P.k * P.l, P.secret_variance, g).reshape(D, P.k, P.l)          # encryption.rs:135-142
    e1 = sample_uniform_coefficients(P.error_bound_1, D * P.k * P.l, g).reshape(D, P.k, P.l)  # :161-167
    e2 = sample_uniform_coefficients(P.error_bound_2, D * P.n * P.l, g).reshape(D, P.n, P.l)  # :196
    return r, e1, e2


def _encrypt_many(all_scalars: np.ndarray, global_pk: GlobalPublicKey, randomness=None) -> List[PvwCiphertext]:
    P = global_pk.params
    D = all_scalars.shape[0]
    if not global_pk.is_full():                                                           # encryption.rs:117-121
        raise PvwError("InvalidParameters", "Global public key is not complete (missing party keys)")
    if not P.verify_correctness_condition():                                              # encryption.rs:124-128
        raise PvwError("InvalidParameters", "Parameters do not satisfy correctness condition - decryption may fail")
    r, e1, e2 = randomness if randomness is not None else _sample_randomness(P, D)
    r, e1, e2 = (np.asarray(x, dtype=np.int64) for x in (r, e1, e2))
    for name, a, shape in (("r", r, (D, P.k, P.l)), ("e1", e1, (D, P.k, P.l)), ("e2", e2, (D, P.n, P.l))):
        if a.shape != shape:
            raise PvwError("DimensionMismatch", f"{name}: expected {shape}, got {a.shape}")
    out: List[PvwCiphertext] = []
    cap = global_pk.engine.capacity
    with global_pk._lock:
        for d0 in range(0, D, cap):
            d1 = min(D, d0 + cap)
            cts = [PvwCiphertext(global_pk) for _ in range(d1 - d0)]
            slots = global_pk._slots.take_many(cts)
            global_pk.engine.encrypt_batch(slots[0], all_scalars[d0:d1], r[d0:d1], e1[d0:d1], e2[d0:d1])
            for ct, s in zip(cts, slots):
                ct._slot = s
            out.extend(cts)
    return out


def _as_scalars(x, what: str) -> np.ndarray:
    """&[u64] of the reference: anything outside [0, 2^64) is not representable there and is rejected, not wrapped"""
    try:
        vals = [int(v) for v in x]
    except (TypeError, ValueError):
        raise PvwError("InvalidParameters", f"{what} must be a sequence of u64")
    for v in vals:
        if not 0 <= v <= _U64:
            raise PvwError("InvalidParameters", f"{what}: {v} does not fit u64")
    return np.array(vals, dtype=np.uint64)


def encrypt(scalars: Sequence[int], global_pk: GlobalPublicKey, randomness=None) -> PvwCiphertext:
    """encryption.rs:105-214.  randomness = (r [k][l], e1 [k][l], e2 [n][l]) or None to sample on the host."""
    P = global_pk.params
    if len(scalars) != P.n:                                                               # :109-115
        raise PvwError("InvalidParameters", f"Must provide exactly n={P.n} scalars, got {len(scalars)}")
    rnd = None if randomness is None else tuple(np.asarray(x, np.int64)[None] for x in randomness)
    ct = _encrypt_many(_as_scalars(scalars, "scalars")[None], global_pk, rnd)[0]
    ct.validate()                                                                          # :211
    return ct


def encrypt_party_shares(party_shares: Sequence[int], party_index: int, global_pk: GlobalPublicKey, randomness=None) -> PvwCiphertext:
    """encryption.rs:221-245"""
    P = global_pk.params
    if party_index >= P.n:
        raise PvwError("InvalidParameters", f"Party index {party_index} exceeds maximum {P.n - 1}")
    if len(party_shares) != P.n:
        raise PvwError("InvalidParameters", f"Party must provide {P.n} shares, got {len(party_shares)}")
    return encrypt(party_shares, global_pk, randomness)


def encrypt_all_party_shares(all_shares: Sequence[Sequence[int]], global_pk: GlobalPublicKey, randomness=None) -> List[PvwCiphertext]:
    """encryption.rs:253-286: D = n dealers in one batched device call (the reference fans out with rayon)."""
    P = global_pk.params
    if len(all_shares) != P.n:
        raise PvwError("InvalidParameters", f"Must provide shares for all {P.n} parties")
    for d, s in enumerate(all_shares):
        if len(s) != P.n:
            raise PvwError("InvalidParameters", f"Dealer {d} provided {len(s)} shares but needs {P.n}")
    m = np.stack([_as_scalars(s, "shares") for s in all_shares])
    return _encrypt_many(m, global_pk, randomness)


def encrypt_broadcast(scalar: int, global_pk: GlobalPublicKey, randomness=None) -> PvwCiphertext:
    """encryption.rs:292-296"""
    return encrypt([scalar] * global_pk.params.n, global_pk, randomness)


# --------------------------------------------------------------------------------------------------------------
# decryption -- src/crypto/decryption.rs
# --------------------------------------------------------------------------------------------------------------
def decrypt_party_value(ciphertext: PvwCiphertext, secret_key: SecretKey, party_index: int) -> int:
    """decryption.rs:249-278.  The reference indexes c2[party_index] unchecked (a panic); here that is IndexOutOfBounds."""
    P = ciphertext.params
    if not 0 <= party_index < P.n:
        raise PvwError("IndexOutOfBounds", f"party index {party_index} out of range for c2 of length {P.n}")
    pk = ciphertext._pk
    with pk._lock:
        slot = ciphertext._resident_slot()
        out = pk.engine.decrypt_batch([party_index], secret_key.secret_coeffs[None], dealer_slots=[slot])
    return int(out[0, 0])


def decrypt_party_shares(all_ciphertexts: Sequence[PvwCiphertext], secret_key: SecretKey, party_index: int) -> List[int]:
    """decryption.rs:281-325: exactly n ciphertexts, one batched device call over the dealers."""
    if len(all_ciphertexts) == 0:
        raise PvwError("InvalidParameters", "No ciphertexts provided")
    P = all_ciphertexts[0].params
    if len(all_ciphertexts) != P.n:
        raise PvwError("InvalidParameters", f"Expected {P.n} ciphertexts, got {len(all_ciphertexts)}")
    if party_index >= P.n:
        raise PvwError("InvalidParameters", f"Party index {party_index} exceeds maximum {P.n - 1}")
    for d, ct in enumerate(all_ciphertexts):
        try:
            ct.validate()
        except PvwError as e:
            raise PvwError("InvalidParameters", f"Ciphertext {d} invalid: {e}")
    pk = all_ciphertexts[0]._pk
    with pk._lock:
        if len(all_ciphertexts) <= pk.engine.capacity:
            # making a spilled ciphertext resident may evict another one: never one of THIS list (its recorded slot would then hold
            # a different ciphertext and a dealer would silently be decrypted from the wrong data)
            slots, pinned = [], set()
            for ct in all_ciphertexts:
                s = ct._resident_slot(pinned)
                pinned.add(s)
                slots.append(s)
            assert all(ct._slot == s for ct, s in zip(all_ciphertexts, slots))
            out = pk.engine.decrypt_batch([party_index], secret_key.secret_coeffs[None], dealer_slots=slots)
            return [int(v) for v in out[0]]
    return [decrypt_party_value(ct, secret_key, party_index) for ct in all_ciphertexts]
