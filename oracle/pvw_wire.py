"""CPU oracle for the pvw-rs wire format: bincode 1.3 encodings of the crate's serde impls.

TEST INFRASTRUCTURE ONLY (see pvw_oracle.py): imported by tests/ and tools that check the CUDA serialisers, never by
the product path.

PARITY UNPINNED.  The byte layout is produced by three third-party crates that are absent from /root/reference and
from this image: serde + bincode 1.3 (`bincode::serialize` = fixed-width little-endian integers, u64 length prefixes,
usize as u64), fhe-math 0.1.0-beta.7 `Poly::to_bytes` (prost encoding of the `Rq` message of
crates/fhe-math/src/proto/rq.proto) and fhe-util `transcode_to_bytes` (LSB-first bit packing).  They are restated here
from their published behaviour; the reference's own tests (tests/serialization.rs) only check round trips and
determinism, no byte-level known answers.  What IS pinned by the reference: the field order and types of every struct
(the hand-written Serialize impls cited below).

    message Rq {                                  // recalled
      enum Representation { UNKNOWN = 0; POWERBASIS = 1; NTT = 2; NTTSHOUP = 3; }
      Representation representation = 1;          // varint, omitted when 0
      uint32 degree = 2;                          // varint
      bytes coefficients = 3;                     // for each modulus q_j in order: the ell residues packed with
                                                  // nbits_j = bit_length(q_j - 1) bits each, LSB first
      bool allow_variable_time = 4;               // false in pvw-rs (parameters.rs:464) -> omitted
    }
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple

from pvw_oracle import Params, Poly, PvwError

REP_POWERBASIS, REP_NTT = 1, 2


# --------------------------------------------------------------------------------------------------------------------
# bincode 1.3 (default options: little endian, fixed-width ints, u64 lengths)
# --------------------------------------------------------------------------------------------------------------------
def bc_u64(x: int) -> bytes:
    return struct.pack("<Q", x)


def bc_i64(x: int) -> bytes:
    return struct.pack("<q", x)


def bc_f32(x: float) -> bytes:
    return struct.pack("<f", x)


def bc_bytes(b: bytes) -> bytes:           # Vec<u8> / String
    return bc_u64(len(b)) + b


def bc_seq(items: Sequence[bytes]) -> bytes:  # Vec<T> of already-encoded items
    return bc_u64(len(items)) + b"".join(items)


class Reader:
    def __init__(self, data: bytes):
        self.d, self.o = memoryview(data), 0

    def take(self, n: int) -> bytes:
        if self.o + n > len(self.d):
            raise PvwError("SerializationError", "unexpected end of input")      # bincode: io::ErrorKind::UnexpectedEof
        b = bytes(self.d[self.o:self.o + n])
        self.o += n
        return b

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]

    def i64(self) -> int:
        return struct.unpack("<q", self.take(8))[0]

    def f32(self) -> float:
        return struct.unpack("<f", self.take(4))[0]

    def bytes_(self) -> bytes:
        return self.take(self.u64())

    def done(self):
        return self.o == len(self.d)


# --------------------------------------------------------------------------------------------------------------------
# fhe-util transcode_to_bytes / transcode_from_bytes, fhe-math Modulus::serialize_vec
# --------------------------------------------------------------------------------------------------------------------
def nbits_of(q: int) -> int:
    return (q - 1).bit_length()


def transcode_to_bytes(a: Sequence[int], nbits: int) -> bytes:
    acc = 0
    for i, v in enumerate(a):
        acc |= (v & ((1 << nbits) - 1)) << (i * nbits)
    return acc.to_bytes((len(a) * nbits + 7) // 8, "little")


def transcode_from_bytes(b: bytes, nbits: int, count: int) -> List[int]:
    acc = int.from_bytes(b, "little")
    return [(acc >> (i * nbits)) & ((1 << nbits) - 1) for i in range(count)]


def varint(x: int) -> bytes:
    out = bytearray()
    while True:
        if x < 0x80:
            out.append(x)
            return bytes(out)
        out.append((x & 0x7F) | 0x80)
        x >>= 7


def read_varint(b: bytes, o: int) -> Tuple[int, int]:
    x, s = 0, 0
    while True:
        if o >= len(b):
            raise PvwError("SerializationError", "truncated varint")
        c = b[o]
        o += 1
        x |= (c & 0x7F) << s
        if c < 0x80:
            return x, o
        s += 7
        if s > 63:
            raise PvwError("SerializationError", "varint too long")


# --------------------------------------------------------------------------------------------------------------------
# Poly::to_bytes / Poly::from_bytes (fhe-math rq/convert.rs + rq/serialize.rs, recalled)
# --------------------------------------------------------------------------------------------------------------------
def poly_to_bytes(P: Params, p: Poly, representation: int = REP_NTT) -> bytes:
    packed = b"".join(transcode_to_bytes(row, nbits_of(q)) for row, q in zip(p, P.moduli))
    out = b""
    if representation:
        out += b"\x08" + varint(representation)
    out += b"\x10" + varint(P.l)
    out += b"\x1a" + varint(len(packed)) + packed      # prost omits empty bytes fields; never empty here
    return out


def poly_from_bytes(P: Params, b: bytes) -> Tuple[Poly, int]:
    """canonical encodings only (field order 1,2,3; no unknown fields) -- returns (residues, representation)"""
    o, rep, degree, packed = 0, 0, 0, b""
    while o < len(b):
        tag, o = read_varint(b, o)
        if tag == 0x08:
            rep, o = read_varint(b, o)
        elif tag == 0x10:
            degree, o = read_varint(b, o)
        elif tag == 0x1A:
            n, o = read_varint(b, o)
            if o + n > len(b):
                raise PvwError("SerializationError", "truncated coefficients")
            packed, o = b[o:o + n], o + n
        elif tag == 0x20:
            _, o = read_varint(b, o)
        else:
            raise PvwError("SerializationError", f"unexpected protobuf tag {tag:#x}")
    if rep not in (REP_POWERBASIS, REP_NTT):
        raise PvwError("SerializationError", "Invalid representation")
    if degree != P.l:
        raise PvwError("SerializationError", "Invalid degree")
    sizes = [(P.l * nbits_of(q) + 7) // 8 for q in P.moduli]
    if len(packed) != sum(sizes):
        raise PvwError("SerializationError", "Invalid coefficients")
    rows, o = [], 0
    for q, s in zip(P.moduli, sizes):
        row = transcode_from_bytes(packed[o:o + s], nbits_of(q), P.l)
        if any(v >= q for v in row):
            raise PvwError("SerializationError", "coefficient not reduced")
        rows.append(row)
        o += s
    return rows, rep


def poly_record_bytes(P: Params) -> int:
    """size of one `Vec<u8>` element holding a polynomial: u64 length + Rq message"""
    return 8 + len(poly_to_bytes(P, [[0] * P.l for _ in P.moduli]))


# --------------------------------------------------------------------------------------------------------------------
# the crate's structs
# --------------------------------------------------------------------------------------------------------------------
def _f32_repr(x: float) -> float:
    return struct.unpack("<f", struct.pack("<f", x))[0]


def params_to_bytes(P: Params) -> bytes:
    """impl Serialize for PvwParameters, src/params/parameters.rs:606-623"""
    return (bc_u64(P.n) + bc_u64(P.k) + bc_u64(P.l) + bc_seq([bc_u64(q) for q in P.moduli]) + bc_f32(P.secret_variance)
            + bc_bytes(str(P.error_bound_1).encode()) + bc_bytes(str(P.error_bound_2).encode()))


def params_read(r: Reader, psi=None) -> Params:
    """impl Deserialize for PvwParameters, parameters.rs:625-664: rebuilt through the builder (so it re-validates)"""
    n, k, l = r.u64(), r.u64(), r.u64()
    moduli = [r.u64() for _ in range(r.u64())]
    var = r.f32()
    try:
        b1, b2 = int(r.bytes_().decode()), int(r.bytes_().decode())
    except ValueError as e:
        raise PvwError("SerializationError", str(e))
    return Params(n, k, l, moduli, var, b1, b2, psi=psi)


def params_from_bytes(b: bytes, psi=None) -> Params:
    return params_read(Reader(b), psi)      # bincode::deserialize tolerates trailing bytes


def same_params(a: Params, b: Params) -> bool:
    return (a.n, a.k, a.l, list(a.moduli), _f32_repr(a.secret_variance), a.error_bound_1, a.error_bound_2) == \
           (b.n, b.k, b.l, list(b.moduli), _f32_repr(b.secret_variance), b.error_bound_1, b.error_bound_2)


def _polys(P: Params, polys: Sequence[Poly]) -> bytes:        # Vec<Vec<u8>>
    return bc_seq([bc_bytes(poly_to_bytes(P, p)) for p in polys])


def _read_polys(P: Params, r: Reader) -> List[Poly]:
    return [poly_from_bytes(P, r.bytes_())[0] for _ in range(r.u64())]


def secret_key_to_bytes(P: Params, coeffs: Sequence[Sequence[int]]) -> bytes:
    """impl Serialize for SecretKey, src/keys/secret_key.rs:294-307: Vec<Vec<i64>> + params"""
    return bc_seq([bc_seq([bc_i64(c) for c in row]) for row in coeffs]) + params_to_bytes(P)


def secret_key_from_bytes(b: bytes, psi=None):
    r = Reader(b)
    coeffs = [[r.i64() for _ in range(r.u64())] for _ in range(r.u64())]
    P = params_read(r, psi)
    # SecretKey::from_coefficients (secret_key.rs:62-88): k rows of l coefficients
    if len(coeffs) != P.k or any(len(row) != P.l for row in coeffs):
        raise PvwError("InvalidParameters", "secret key shape")
    return P, coeffs


def public_key_to_bytes(P: Params, key_polys: Sequence[Poly]) -> bytes:
    """impl Serialize for PublicKey, src/keys/public_key.rs:471-487"""
    return _polys(P, key_polys) + params_to_bytes(P)


def public_key_from_bytes(b: bytes, psi=None):
    r = Reader(b)
    # the polynomial records precede the parameters they need: find the parameters first (the deserialiser does the
    # same by materialising Vec<Vec<u8>> before building the context, public_key.rs:496-519)
    raw = [r.bytes_() for _ in range(r.u64())]
    P = params_read(r, psi)
    return P, [poly_from_bytes(P, x)[0] for x in raw]


def crs_to_bytes(P: Params, A: Sequence[Sequence[Poly]]) -> bytes:
    """impl Serialize for PvwCrs, src/params/crs.rs:228-249: Vec<Vec<Vec<u8>>> (row-major) + params"""
    return bc_seq([_polys(P, row) for row in A]) + params_to_bytes(P)


def crs_from_bytes(b: bytes, psi=None):
    r = Reader(b)
    P, A = _crs_read(r, psi)
    return P, A


def _crs_read(r: Reader, psi=None):
    raw = [[r.bytes_() for _ in range(r.u64())] for _ in range(r.u64())]
    P = params_read(r, psi)
    cols = len(raw[0]) if raw else 0
    if any(len(row) != cols for row in raw):                       # Array2::from_shape_vec failure, crs.rs:287-288
        raise PvwError("SerializationError", "ragged matrix")
    return P, [[poly_from_bytes(P, x)[0] for x in row] for row in raw]


def global_public_key_to_bytes(P: Params, B, A, num_keys: int, error_polys) -> bytes:
    """impl Serialize for GlobalPublicKey, public_key.rs:522-552: matrix, crs, num_keys, params, error_polynomials"""
    return (bc_seq([_polys(P, row) for row in B]) + crs_to_bytes(P, A) + bc_u64(num_keys) + params_to_bytes(P)
            + bc_seq([_polys(P, row) for row in error_polys]))


def global_public_key_from_bytes(b: bytes, psi=None):
    r = Reader(b)
    raw = [[r.bytes_() for _ in range(r.u64())] for _ in range(r.u64())]
    _, A = _crs_read(r, psi)
    num_keys = r.u64()
    P = params_read(r, psi)
    err_raw = [[r.bytes_() for _ in range(r.u64())] for _ in range(r.u64())]
    B = [[poly_from_bytes(P, x)[0] for x in row] for row in raw]
    errs = [[poly_from_bytes(P, x)[0] for x in row] for row in err_raw]
    return P, B, A, num_keys, errs


def ciphertext_to_bytes(P: Params, c1: Sequence[Poly], c2: Sequence[Poly]) -> bytes:
    """impl Serialize for PvwCiphertext, src/crypto/encryption.rs:298-317"""
    return _polys(P, c1) + _polys(P, c2) + params_to_bytes(P)


def ciphertext_from_bytes(b: bytes, psi=None):
    """impl Deserialize for PvwCiphertext, encryption.rs:319-354 (no length validation there; validate() is separate)"""
    r = Reader(b)
    raw1 = [r.bytes_() for _ in range(r.u64())]
    raw2 = [r.bytes_() for _ in range(r.u64())]
    P = params_read(r, psi)
    return P, [poly_from_bytes(P, x)[0] for x in raw1], [poly_from_bytes(P, x)[0] for x in raw2]
