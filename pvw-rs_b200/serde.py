"""The crate's `serde` feature (Cargo.toml:37-39): `bincode::serialize(&x)` / `bincode::deserialize(&bytes)` of every
public type, byte-compatible with the hand-written impls
  PvwParameters   src/params/parameters.rs:606-664        SecretKey        src/keys/secret_key.rs:294-326
  PublicKey       src/keys/public_key.rs:471-519          GlobalPublicKey  src/keys/public_key.rs:522-622
  PvwCrs          src/params/crs.rs:228-295               PvwCiphertext    src/crypto/encryption.rs:298-354
Polynomials are bit-packed into / parsed from their `Poly::to_bytes` records by the CUDA kernels of csrc/wire.cu where they
live (device store, B, A); this module only frames the fixed-size pieces.  The third-party encodings are recalled
(bincode 1.3, prost, fhe-util transcode): parity unpinned, see DESIGN.md.

    blob = serde.serialize(ciphertext)                      # bincode::serialize(&ciphertext)
    ct   = serde.deserialize(PvwCiphertext, blob, gpk)      # bincode::deserialize::<PvwCiphertext>(&blob)

Deserialisers that the reference gives a fresh `Arc<PvwParameters>` built from the embedded parameters take the owning
object instead (`global_pk` for ciphertexts) or build new parameters when none is given.
"""
from __future__ import annotations

import struct
from typing import Optional

import numpy as np

from .api import GlobalPublicKey, PublicKey, PvwCiphertext, PvwCrs, PvwParameters, SecretKey
from .errors import PvwError


def _u64(x: int) -> bytes:
    return struct.pack("<Q", x)


class _Reader:
    def __init__(self, data):
        self.d = memoryview(data).cast("B")
        self.o = 0

    def take(self, n: int):
        if self.o + n > len(self.d):
            raise PvwError("InsufficientData", f"expected {self.o + n} bytes, got {len(self.d)}")
        v = self.d[self.o:self.o + n]
        self.o += n
        return v

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]


# ---- PvwParameters ---------------------------------------------------------------------------------------------
def _params_read(r: _Reader, psi=None, device: int = 0) -> PvwParameters:
    n, k, l, L = r.u64(), r.u64(), r.u64(), r.u64()
    if L > 1 << 16:
        raise PvwError("DeserializationError", "implausible modulus count")
    moduli = list(struct.unpack(f"<{L}Q", r.take(8 * L)))
    (var,) = struct.unpack("<f", r.take(4))
    try:
        b1 = int(bytes(r.take(r.u64())).decode())
        b2 = int(bytes(r.take(r.u64())).decode())
    except (ValueError, UnicodeDecodeError) as e:
        raise PvwError("DeserializationError", str(e))
    # parameters.rs:652-661: rebuilt through the builder, so every validation runs again
    return PvwParameters(n, k, l, moduli, var, b1, b2, psi=psi, device=device)


def _same_params(a: PvwParameters, blob) -> bool:
    return bytes(blob) == a._probe.wire_params()


def _probe_record_bytes(data, off: int) -> int:
    """size of the polynomial record (u64 length + Rq message) that starts at `off`"""
    r = _Reader(data)
    r.o = off
    return 8 + r.u64()


# ---- the public entry points -------------------------------------------------------------------------------------
def serialize(x) -> bytes:
    """bincode::serialize(&x)"""
    if isinstance(x, PvwParameters):
        return x._probe.wire_params()
    if isinstance(x, SecretKey):                                             # Vec<Vec<i64>> + params
        c = np.ascontiguousarray(x.secret_coeffs, dtype="<i8")
        rows = b"".join(_u64(c.shape[1]) + row.tobytes() for row in c)
        return _u64(c.shape[0]) + rows + serialize(x.params)
    if isinstance(x, PublicKey):                                             # Vec<Vec<u8>> + params
        eng = x.params._probe
        return _u64(len(x.key_polynomials)) + eng.wire_polys_serialize(x.key_polynomials) + serialize(x.params)
    if isinstance(x, PvwCrs):
        return x._eng().wire_crs_serialize()
    if isinstance(x, GlobalPublicKey):                                       # matrix, crs, num_keys, params, error_polynomials
        eng, P = x.engine, x.params
        errs = b"".join(_u64(len(e)) + (eng.wire_polys_serialize(e) if len(e) else b"") for e in x.error_polynomials)
        return (_u64(P.n) + eng.wire_pk_serialize_rows(0, P.n).tobytes() + eng.wire_crs_serialize() + _u64(x.num_keys) + serialize(P)
                + _u64(len(x.error_polynomials)) + errs)
    if isinstance(x, PvwCiphertext):
        with x._pk._lock:
            return x._pk.engine.wire_ct_serialize(x._resident_slot(), 1)[0].tobytes()
    raise TypeError(f"cannot serialize {type(x).__name__}")


def serialize_ciphertexts(cts) -> np.ndarray:
    """many ciphertexts of one key in one device pass when their slots are consecutive -> uint8 [len(cts)][ciphertext_bytes]"""
    if not cts:
        return np.zeros((0, 0), dtype=np.uint8)
    pk = cts[0]._pk
    with pk._lock:
        slots = [c._resident_slot() for c in cts]
        if slots == list(range(slots[0], slots[0] + len(slots))):
            return pk.engine.wire_ct_serialize(slots[0], len(slots))
        return np.stack([pk.engine.wire_ct_serialize(s, 1)[0] for s in slots])


def deserialize(cls, data, owner=None, psi=None, device: int = 0):
    """bincode::deserialize::<cls>(&data).  `owner`: the GlobalPublicKey a ciphertext belongs to / the PvwParameters to
    check the embedded ones against (built from the blob when omitted)."""
    r = _Reader(data)
    if cls is PvwParameters:
        return _params_read(r, psi, device)
    if cls is SecretKey:
        rows = r.u64()
        coeffs = []
        for _ in range(rows):
            cnt = r.u64()
            coeffs.append(np.frombuffer(r.take(8 * cnt), dtype="<i8"))
        params = _params_read(r, psi, device) if owner is None else owner
        if owner is not None and not _same_params(owner, r.take(len(owner._probe.wire_params()))):
            raise PvwError("DeserializationError", "embedded parameters differ")
        if len({len(c) for c in coeffs}) > 1:
            raise PvwError("InvalidParameters", "ragged secret key coefficients")
        return SecretKey(params, np.stack(coeffs) if coeffs else np.zeros((0, 0), np.int64))    # from_coefficients validates the shape
    if cls is PublicKey:
        k = r.u64()
        rec = _probe_record_bytes(data, 8) if k else 0
        body = r.take(k * rec)
        params = _params_read(r, psi, device) if owner is None else owner
        if owner is not None and not _same_params(owner, r.take(len(owner._probe.wire_params()))):
            raise PvwError("DeserializationError", "embedded parameters differ")
        if rec != params._probe.wire_layout.record_bytes:
            raise PvwError("DeserializationError", "polynomial record size does not match the parameters")
        return PublicKey(params, params._probe.wire_polys_deserialize(body, k))
    if cls is PvwCrs:
        k = r.u64()
        k2 = r.u64() if k else 0
        if k == 0 or k2 != k:
            raise PvwError("DeserializationError", "CRS matrix is not square")
        rec = _probe_record_bytes(data, 16)
        r.o = 8 + k * (8 + k * rec)
        params = _params_read(r, psi, device) if owner is None else owner
        crs = PvwCrs.__new__(PvwCrs)
        crs.params, crs._engine = params, params.new_engine(0, 1)
        crs._engine.wire_crs_deserialize(data)                               # validates every record and the embedded parameters
        crs.matrix = crs._engine.crs_download()
        return crs
    if cls is GlobalPublicKey:
        n = r.u64()
        k = r.u64() if n else 0
        if n == 0 or k == 0:
            raise PvwError("DeserializationError", "empty public key matrix")
        rec = _probe_record_bytes(data, 16)
        row = 8 + k * rec
        rows = r.d[8:8 + n * row]
        if len(rows) != n * row:
            raise PvwError("InsufficientData", f"expected {8 + n * row} bytes, got {len(r.d)}")
        r.o = 8 + n * row
        crs_off = r.o
        crs_len = 8 + k * row
        r.o += crs_len
        P0 = _params_read(r, psi, device)                                    # the CRS's own copy of the parameters
        crs = deserialize(PvwCrs, r.d[crs_off:r.o], owner=P0)
        num_keys = r.u64()
        tail = r.take(len(P0._probe.wire_params()))
        if not _same_params(P0, tail):
            raise PvwError("DeserializationError", "embedded parameters differ")
        gpk = GlobalPublicKey(crs)
        if n != P0.n:
            raise PvwError("DeserializationError", "matrix rows != n")
        if num_keys > n:
            raise PvwError("DeserializationError", "num_keys exceeds n")
        # rows at and above num_keys are the zero polynomials GlobalPublicKey::new filled in (public_key.rs:196-208)
        gpk.engine.wire_pk_deserialize_rows(0, num_keys, rows[:num_keys * row])
        for _ in range(r.u64()):
            cnt = r.u64()
            gpk.error_polynomials.append(gpk.engine.wire_polys_deserialize(r.take(cnt * rec), cnt) if cnt
                                         else np.zeros((0, P0.L, P0.l), dtype=np.uint64))
        return gpk
    if cls is PvwCiphertext:
        if not isinstance(owner, GlobalPublicKey):
            raise PvwError("InvalidParameters", "deserialize(PvwCiphertext, data, owner=<GlobalPublicKey>)")
        ct = PvwCiphertext(owner)
        with owner._lock:
            ct._slot = owner._slots.take(ct)
            try:
                owner.engine.wire_ct_deserialize(ct._slot, 1, data)
            except PvwError:
                owner._slots._release(ct._slot)
                ct._slot = None
                raise
        return ct
    raise TypeError(f"cannot deserialize {cls!r}")
