"""CPU oracle for the pvw-rs hot path (multi-receiver PVW encrypt + per-party decrypt).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and there only as the checker (or the timed CPU baseline), never as the thing shipped.

PARITY UNPINNED at the fhe-math boundary.  The reference (gnosisguild/pvw-rs, /root/reference) is a Rust
crate whose ring arithmetic lives in un-vendored git dependencies that are absent here:
  fhe-math / fhe-util / fhe-traits 0.1.0-beta.7 @ gnosisguild/fhe.rs#364335035e3a801573539274b2d3052d5b69098a
  (Cargo.lock:294-327), num-bigint 0.4.6, rand 0.8.5, rand_chacha 0.3.1.
No Rust toolchain exists in this image, and the reference holds no known-answer vectors for ciphertext
residues, NTT outputs or the primitive root psi.  This file therefore restates the *published* algorithms
(negacyclic NTT with bit-reversed output, canonical residues, CRT lift, truncated BigInt division) and is
pinned only by (i) the reference's own behavioural tests restated in tests/ (gadget structure, RNS reduce /
lift round trips, the rounding rule of the Delta division, end-to-end recovery of m), and (ii)
psi-independent algebra (schoolbook negacyclic product).  psi is an explicit parameter everywhere.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Exact integer arithmetic throughout (Python ints); meant for small cases and for generating golden vectors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

MASK64 = (1 << 64) - 1

# --------------------------------------------------------------------------------------------------
# Deterministic synthetic streams.  Stand-ins for thread_rng(): same distributions as
# src/sampling/uniform.rs:5-70, not the same streams.
# --------------------------------------------------------------------------------------------------
DEFAULT_SEED = 0x5056572D42323030

TAG_A, TAG_SK, TAG_KE, TAG_R, TAG_E1, TAG_E2, TAG_M, TAG_B = 1, 2, 3, 4, 5, 6, 7, 8


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def stream_u64(seed: int, tag: int, index: int) -> int:
    return splitmix64((seed ^ (tag << 56) ^ index) & MASK64)


def uniform_residue(u: int, q: int) -> int:
    return (u * q) >> 64


def uniform_sym(u: int, b: int) -> int:
    """uniform integer in [-b, b] (shape of sample_uniform_coefficients, uniform.rs:5-22)"""
    return ((u * (2 * b + 1)) >> 64) - b


def cbd(u: int, variance: float) -> int:
    """centred binomial sample (shape of sample_vec_cbd, uniform.rs:27-70)"""
    if abs(variance - 0.5) < 1.2e-7:
        return (u & 1) - ((u >> 1) & 1)
    v = int(variance)
    mask = (1 << (2 * v)) - 1
    return bin(u & mask).count("1") - bin((u >> (2 * v)) & mask).count("1")


# --------------------------------------------------------------------------------------------------
# fhe-math default primitive root (recalled from fhe.rs 0.1.0-beta.7 ntt/native.rs `primitive_root`;
# source NOT available here -- unverified).  ChaCha8Rng::seed_from_u64(0); up to 100 tries of
# root = gen_range(0..p) ^ ((p-1)/2n); accept when root^(2n) == 1 and root^n != 1.
# --------------------------------------------------------------------------------------------------
def _rotl32(x, n):
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


def _chacha_block(key_words, counter, rounds=8):
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [
        counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, 0, 0]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl32(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl32(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl32(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl32(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & 0xFFFFFFFF for i in range(16)]


class ChaCha8Rng:
    """rand_chacha 0.3.1 ChaCha8Rng word stream (64-bit block counter, zero stream id)."""

    def __init__(self, seed32: bytes):
        assert len(seed32) == 32
        self.key = [int.from_bytes(seed32[4 * i:4 * i + 4], "little") for i in range(8)]
        self.counter = 0
        self.buf: List[int] = []

    @classmethod
    def seed_from_u64(cls, state: int) -> "ChaCha8Rng":
        # rand_core 0.6 SeedableRng::seed_from_u64: PCG32 expansion
        MUL, INC = 6364136223846793005, 11634580027462260723
        out = b""
        for _ in range(8):
            state = (state * MUL + INC) & MASK64
            xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
            rot = state >> 59
            x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF
            out += x.to_bytes(4, "little")
        return cls(out)

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = _chacha_block(self.key, self.counter)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return (hi << 32) | lo

    def gen_range_u64(self, high: int) -> int:
        """rand 0.8.5 UniformInt::<u64>::sample_single(0, high): widening-multiply rejection."""
        rng_range = high
        lz = 64 - rng_range.bit_length()
        zone = (((rng_range << lz) & MASK64) - 1) & MASK64
        while True:
            v = self.next_u64()
            m = v * rng_range
            hi, lo = m >> 64, m & MASK64
            if lo <= zone:
                return hi


def fhe_math_default_psi(q: int, ell: int) -> int:
    lam = (q - 1) // (2 * ell)
    rng = ChaCha8Rng.seed_from_u64(0)
    for _ in range(100):
        root = rng.gen_range_u64(q)
        root = pow(root, lam, q)
        if pow(root, 2 * ell, q) == 1 and pow(root, ell, q) != 1:
            return root
    raise ValueError("no primitive root found")


def is_primitive_2l_root(psi: int, q: int, ell: int) -> bool:
    return pow(psi, ell, q) == q - 1


# --------------------------------------------------------------------------------------------------
# helpers mirroring num-bigint semantics
# --------------------------------------------------------------------------------------------------
def tdiv(a: int, b: int) -> int:
    """Rust BigInt `/`: truncated toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def trem(a: int, b: int) -> int:
    """Rust BigInt `%`: remainder with the sign of the dividend."""
    return a - b * tdiv(a, b)


def big_to_f64(x: int) -> float:
    """num-bigint 0.4 ToPrimitive::to_f64: correctly rounded, +inf on overflow (recalled)."""
    try:
        return float(x)
    except OverflowError:
        return math.inf if x > 0 else -math.inf


def iroot(x: int, n: int) -> int:
    """floor(x ** (1/n)) -- BigUint::nth_root (parameters.rs:156)."""
    if x < 2:
        return x
    lo, hi = 1, 1 << (x.bit_length() // n + 1)
    while lo < hi:
        mid = (lo + hi + 1) >> 1
        if mid ** n <= x:
            lo = mid
        else:
            hi = mid - 1
    return lo


def brv(i: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def is_prime(n: int) -> bool:
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


class PvwError(Exception):
    """errors.rs:11-73 -- only the variant name and message matter here."""

    def __init__(self, variant: str, msg: str = ""):
        super().__init__(f"{variant}: {msg}")
        self.variant = variant


Poly = List[List[int]]  # [L][ell] canonical residues; representation tracked by the caller


# --------------------------------------------------------------------------------------------------
# Parameters (src/params/parameters.rs)
# --------------------------------------------------------------------------------------------------
@dataclass
class Params:
    n: int
    k: int
    l: int
    moduli: Sequence[int]
    secret_variance: float = 0.5          # parameters.rs:166
    error_bound_1: int = 100              # parameters.rs:167
    error_bound_2: int = 200              # parameters.rs:168
    psi: Optional[Sequence[int]] = None   # explicit 2l-th primitive roots; None -> fhe-math default (recalled)
    Q: int = field(init=False)
    delta: int = field(init=False)
    delta_power_l_minus_1: int = field(init=False)
    t: int = field(init=False)

    def __post_init__(self):
        # PvwParametersBuilder::build, parameters.rs:117-195
        if self.n == 0:
            raise PvwError("InvalidParameters", "n must be > 0")
        if self.k == 0:
            raise PvwError("InvalidParameters", "k must be > 0")
        l = self.l
        if l < 8 or (l & (l - 1)) != 0:
            raise PvwError("InvalidParameters", "l must be power of 2 and >= 8 (fhe.rs Context requirement)")
        mods = list(self.moduli)
        # fhe-math Context::new / Modulus::new / supports_ntt (Appendix B of SURVEY.md)
        if len(mods) == 0:
            raise PvwError("InvalidParameters", "Context creation failed: no moduli")
        for q in mods:
            if q < 2 or q >= (1 << 62) or not is_prime(q) or (q - 1) % (2 * l) != 0:
                raise PvwError("InvalidParameters", f"Context creation failed: modulus {q}")
        if len(set(mods)) != len(mods):
            raise PvwError("InvalidParameters", "Context creation failed: repeated modulus")
        self.moduli = mods
        self.Q = 1
        for q in mods:
            self.Q *= q
        self.delta = iroot(self.Q, l)                                  # parameters.rs:156
        self.delta_power_l_minus_1 = self.delta ** (l - 1)             # parameters.rs:159-163
        self.t = (self.n - 1) // 2                                     # parameters.rs:169
        if self.error_bound_1 <= 0:
            raise PvwError("InvalidParameters", "error_bound_1 must be positive")
        if self.error_bound_2 <= 0:
            raise PvwError("InvalidParameters", "error_bound_2 must be positive")
        if self.psi is None:
            self.psi = [fhe_math_default_psi(q, l) for q in mods]
        self.psi = list(self.psi)
        for q, p in zip(mods, self.psi):
            if not is_primitive_2l_root(p, q, l):
                raise PvwError("InvalidParameters", f"psi {p} is not a primitive 2l-th root mod {q}")
        self._logl = l.bit_length() - 1
        # evaluation points of the forward transform: slot i <-> psi^(2*brv(i)+1)   
        self._pts = [[pow(p, 2 * brv(i, self._logl) + 1, q) for i in range(l)] for q, p in zip(mods, self.psi)]

    @property
    def L(self) -> int:
        return len(self.moduli)

    # -- representation changes (fhe-math Poly::change_representation) ----------------
    def ntt_forward(self, p: Poly) -> Poly:
        """PowerBasis -> Ntt: natural-order input, bit-reversed-order output."""
        out = []
        for j, q in enumerate(self.moduli):
            row = p[j]
            res = []
            for i in range(self.l):
                x = self._pts[j][i]
                acc = 0
                for c in reversed(row):      # Horner
                    acc = (acc * x + c) % q
                res.append(acc)
            out.append(res)
        return out

    def ntt_backward(self, p: Poly) -> Poly:
        """Ntt -> PowerBasis (exact inverse of ntt_forward)."""
        l = self.l
        out = []
        for j, q in enumerate(self.moduli):
            linv = pow(l, -1, q)
            pts_inv = [pow(x, -1, q) for x in self._pts[j]]
            res = []
            for t in range(l):
                acc = 0
                for i in range(l):
                    acc = (acc + p[j][i] * pow(pts_inv[i], t, q)) % q
                res.append(acc * linv % q)
            out.append(res)
        return out

    # -- slot-wise ring ops (fhe-math Poly Mul/Add/Sub/Neg on refs; canonical results) -------------
    def zero(self) -> Poly:
        return [[0] * self.l for _ in self.moduli]

    def mul(self, a: Poly, b: Poly) -> Poly:
        return [[x * y % q for x, y in zip(ra, rb)] for ra, rb, q in zip(a, b, self.moduli)]

    def add(self, a: Poly, b: Poly) -> Poly:
        return [[(x + y) % q for x, y in zip(ra, rb)] for ra, rb, q in zip(a, b, self.moduli)]

    def sub(self, a: Poly, b: Poly) -> Poly:
        return [[(x - y) % q for x, y in zip(ra, rb)] for ra, rb, q in zip(a, b, self.moduli)]

    # -- conversions --------------------------------------------------------------------------------
    def bigints_to_poly(self, bigints: Sequence[int]) -> Poly:
        """parameters.rs:420-474: ((x % q) + q) % q per limb, row = modulus, col = coefficient. PowerBasis."""
        if len(bigints) != self.l:
            raise PvwError("InvalidParameters", f"Expected {self.l} coefficients, got {len(bigints)}")
        out = []
        for q in self.moduli:
            row = []
            for c in bigints:
                r = trem(c, q)            # Rust `%`
                if r < 0:
                    r += q
                row.append(r)
            out.append(row)
        return out

    def from_coefficients(self, coeffs: Sequence[int]) -> Poly:
        """fhe-math Poly::from_coefficients(&[i64]) -- same map (pinned by tests/params.rs:732-767)."""
        return self.bigints_to_poly(list(coeffs))

    def lift(self, p_power: Poly) -> List[int]:
        """Vec<BigUint>::from(&Poly) -- CRT lift of every coefficient to [0, Q)  (RnsContext::lift)."""
        res = []
        for t in range(self.l):
            acc = 0
            for j, q in enumerate(self.moduli):
                qh = self.Q // q
                acc += p_power[j][t] * pow(qh, -1, q) % q * qh
            res.append(acc % self.Q)
        return res

    def gadget_vector(self) -> List[int]:
        """parameters.rs:311-325: [1, D, D^2, ..., D^(l-1)]"""
        return [self.delta ** i for i in range(self.l)]

    def gadget_polynomial(self) -> Poly:
        """parameters.rs:288-308 (Ntt form)."""
        return self.ntt_forward(self.bigints_to_poly(self.gadget_vector()))

    def encode_scalar(self, scalar: int) -> Poly:
        """parameters.rs:346-367: scalar (i64) * [1, D, ..., D^(l-1)] -> RNS -> Ntt."""
        assert -(1 << 63) <= scalar < (1 << 63)
        return self.ntt_forward(self.bigints_to_poly([scalar * g for g in self.gadget_vector()]))

    # -- correctness condition -----------------------------------------------------------------------
    def verify_correctness_condition(self) -> bool:
        """parameters.rs:510-551, f64 arithmetic in the reference's evaluation order."""
        n, k, l = float(self.n), float(self.k), float(self.l)
        b1 = big_to_f64(self.error_bound_1)
        b2 = big_to_f64(self.error_bound_2)
        sqrt_nl = math.sqrt(n * l) if n * l > 0.0 else math.inf
        sqrt_n = math.sqrt(n) if n > 0.0 else math.inf
        first = b2 * sqrt_nl * (1.0 + sqrt_n)
        second = 2.0 * b1 * k * l
        sqrt_nkl = math.sqrt(n * k * l) if n * k * l > 0.0 else math.inf
        third = 14.0 * b1 * sqrt_nkl
        total = first + second + third
        return big_to_f64(self.delta_power_l_minus_1) > total

    @staticmethod
    def suggest_error_bounds(n, k, l, moduli, variance, psi=None):
        """parameters.rs:554-603"""
        tmp = Params(n, k, l, moduli, variance, 1, 1, psi=psi)
        dp = big_to_f64(tmp.delta_power_l_minus_1)
        nf, kf, lf = float(n), float(k), float(l)
        c1 = 2.0 * kf * lf + 14.0 * math.sqrt(nf * kf * lf)
        c2 = math.sqrt(nf * lf) * (1.0 + math.sqrt(nf))
        for b1 in (50, 100, 200, 500, 1000, 2000):
            for b2 in (50, 100, 200, 500, 1000, 2000):
                if dp > float(b1) * c1 + float(b2) * c2:
                    return b1, b2
        raise PvwError("InvalidParameters", "Cannot find suitable error bounds")


# --------------------------------------------------------------------------------------------------
# Key generation (src/params/crs.rs:138-171, src/keys/public_key.rs:111-147)
# --------------------------------------------------------------------------------------------------
def secret_key_polys(P: Params, sk_coeffs: Sequence[Sequence[int]]) -> List[Poly]:
    """SecretKey::get_polynomial for every j (secret_key.rs:98-112): from_coefficients + forward NTT."""
    return [P.ntt_forward(P.from_coefficients(c)) for c in sk_coeffs]


def keygen(P: Params, A: List[List[Poly]], sk_coeffs, e_coeffs) -> List[Poly]:
    """b[c] = sum_j NTT(s[j]) * A[j][c] + NTT(e[c])   (note the transposed index, crs.rs:152-165)."""
    s_hat = secret_key_polys(P, sk_coeffs)
    out = []
    for c in range(P.k):
        acc = P.zero()
        for j in range(P.k):
            acc = P.add(acc, P.mul(s_hat[j], A[j][c]))
        out.append(P.add(acc, P.ntt_forward(P.bigints_to_poly(e_coeffs[c]))))
    return out


# --------------------------------------------------------------------------------------------------
# Encryption with explicit randomness (src/crypto/encryption.rs:105-214 with r, e1, e2 as inputs)
# --------------------------------------------------------------------------------------------------
def multiply_by_randomness(P: Params, A: List[List[Poly]], r_hat: List[Poly]) -> List[Poly]:
    """crs.rs:177-205: out[i] = sum_j A[i][j] * r[j]"""
    if len(r_hat) != P.k:
        raise PvwError("DimensionMismatch", f"expected {P.k}, got {len(r_hat)}")
    out = []
    for i in range(P.k):
        acc = P.zero()
        for j in range(P.k):
            acc = P.add(acc, P.mul(A[i][j], r_hat[j]))
        out.append(acc)
    return out


def encrypt_explicit(P: Params, A, B, scalars, r, e1, e2, num_keys=None):
    """encryption.rs:105-214.  scalars: n u64; r, e1: k x l ints; e2: n x l ints.  Returns (c1, c2) in Ntt form."""
    if len(scalars) != P.n:                                                    # :109
        raise PvwError("InvalidParameters", f"Must provide exactly n={P.n} scalars, got {len(scalars)}")
    if num_keys is not None and num_keys < P.n:                                # :117 is_full()
        raise PvwError("InvalidParameters", "Global public key is not complete (missing party keys)")
    if not P.verify_correctness_condition():                                   # :124
        raise PvwError("InvalidParameters", "Parameters do not satisfy correctness condition - decryption may fail")
    r_hat = [P.ntt_forward(P.from_coefficients(c)) for c in r]                 # :147-154
    c1 = multiply_by_randomness(P, A, r_hat)                                   # :158
    for i in range(P.k):                                                       # :161-173
        c1[i] = P.add(c1[i], P.ntt_forward(P.bigints_to_poly(e1[i])))
    c2 = []
    for p in range(P.n):                                                       # :177-200
        acc = P.zero()
        for j in range(P.k):
            acc = P.add(acc, P.mul(B[p][j], r_hat[j]))
        m = scalars[p] & MASK64
        m_i64 = m - (1 << 64) if m >= (1 << 63) else m                         # `scalars[p] as i64`, :195
        enc = P.encode_scalar(m_i64)
        e2p = P.ntt_forward(P.bigints_to_poly(e2[p]))
        c2.append(P.add(P.add(acc, enc), e2p))                                 # :198
    return c1, c2


# --------------------------------------------------------------------------------------------------
# Decryption (src/crypto/decryption.rs)
# --------------------------------------------------------------------------------------------------
def centre(P: Params, x: int) -> int:
    """center_coefficient_with_precision, decryption.rs:139-152"""
    return x - P.Q if x > P.Q // 2 else x


def _const_poly(P: Params, value: int) -> Poly:
    return P.ntt_forward(P.bigints_to_poly([value] + [0] * (P.l - 1)))


def _extract_coefficient_as_poly(P: Params, poly: Poly, idx: int) -> Poly:
    """decryption.rs:109-137"""
    coeffs = P.lift(P.ntt_backward(poly))
    val = 0 if idx >= len(coeffs) else centre(P, coeffs[idx])
    return _const_poly(P, val)


def _extract_constant_term_bigint(P: Params, poly: Poly) -> int:
    """decryption.rs:209-224"""
    return centre(P, P.lift(P.ntt_backward(poly))[0])


def decode_scalar_pvw_rns(P: Params, noisy: Poly) -> int:
    """Literal, function-by-function restatement of decryption.rs:10-58 and its helpers :61-247."""
    ell = P.l
    delta_poly = _const_poly(P, P.delta)                                        # :14, :61-75
    tmp = []
    for i in range(ell - 1):                                                    # :19-27
        z_i = _extract_coefficient_as_poly(P, noisy, i)
        z_i1 = _extract_coefficient_as_poly(P, noisy, i + 1)
        tmp.append(P.sub(P.mul(z_i, delta_poly), z_i1))
    last = tmp[0]                                                               # :30-33
    for i in range(1, ell - 1):
        last = P.add(P.mul(last, delta_poly), tmp[i])
    delta_power_poly = _const_poly(P, P.delta ** (ell - 1))                     # :36, :78-97
    # reduce_modulo_poly :154-178
    poly_const = _extract_constant_term_bigint(P, last)
    mod_const = _extract_constant_term_bigint(P, delta_power_poly)
    reduced = trem(poly_const, mod_const)
    half_mod = tdiv(mod_const, 2)
    if reduced > half_mod:
        reduced -= mod_const
    elif reduced < -half_mod:
        reduced += mod_const
    tmp.append(_const_poly(P, reduced))
    noise = [None] * ell                                                        # :41-48
    noise[ell - 1] = tmp[ell - 1]
    for i in range(ell - 2, -1, -1):
        numerator = P.sub(noise[i + 1], tmp[i])
        # divide_by_delta_rns :180-207
        pc = _extract_constant_term_bigint(P, numerator)
        dc = _extract_constant_term_bigint(P, delta_poly)
        if dc == 0:
            quo = 0
        elif pc < 0:
            quo = tdiv(2 * pc - dc, 2 * dc)
        else:
            quo = tdiv(2 * pc + dc, 2 * dc)
        noise[i] = _const_poly(P, quo)
    z0 = _extract_coefficient_as_poly(P, noisy, 0)                              # :51-53
    minus_one = _const_poly(P, -1)
    pt_poly = P.sub(P.mul(z0, minus_one), noise[0])
    # extract_constant_term_as_u64 :226-247
    c = _extract_constant_term_bigint(P, pt_poly)
    return _to_u64_rule(P, c)


def _to_u64_rule(P: Params, c: int) -> int:
    """decryption.rs:226-247"""
    if c < 0:
        if -c <= 1000:
            return 0
        pos = trem(c + P.Q, P.Q)
        return pos if pos < (1 << 64) else 0
    return c if c < (1 << 64) else 0


def decode_scalar_fast(P: Params, z: Sequence[int]) -> int:
    """Scalar restatement of decode_scalar_pvw_rns on the lifted coefficients z in [0,Q)^l.
    Every 'constant polynomial' operation of the reference is arithmetic on a scalar mod Q."""
    Q, D, ell = P.Q, P.delta, P.l
    tmp = [(z[i] * D - z[i + 1]) % Q for i in range(ell - 1)]
    last = tmp[0]
    for i in range(1, ell - 1):
        last = (last * D + tmp[i]) % Q
    M = centre(P, (D ** (ell - 1)) % Q)
    red = trem(centre(P, last), M)
    half = tdiv(M, 2)
    if red > half:
        red -= M
    elif red < -half:
        red += M
    noise = red % Q
    Dc = centre(P, D % Q)
    for i in range(ell - 2, -1, -1):
        num = centre(P, (noise - tmp[i]) % Q)
        if Dc == 0:
            quo = 0
        elif num < 0:
            quo = tdiv(2 * num - Dc, 2 * Dc)
        else:
            quo = tdiv(2 * num + Dc, 2 * Dc)
        noise = quo % Q
    pt = centre(P, (-z[0] - noise) % Q)
    return _to_u64_rule(P, pt)


def decrypt_noisy(P: Params, c1: List[Poly], c2_p: Poly, sk_coeffs) -> Poly:
    """decryption.rs:257-274: sum_j NTT(s[j]) * c1[j] - c2[p]   (Ntt form)"""
    s_hat = secret_key_polys(P, sk_coeffs)
    acc = P.zero()
    for j in range(P.k):
        acc = P.add(acc, P.mul(s_hat[j], c1[j]))
    return P.sub(acc, c2_p)


def decrypt_party_value(P: Params, c1, c2, sk_coeffs, party_index: int, literal: bool = False) -> int:
    """decryption.rs:249-278"""
    noisy = decrypt_noisy(P, c1, c2[party_index], sk_coeffs)
    if literal:
        return decode_scalar_pvw_rns(P, noisy)
    return decode_scalar_fast(P, P.lift(P.ntt_backward(noisy)))


def decrypt_party_shares(P: Params, all_cts, sk_coeffs, party_index: int) -> List[int]:
    """decryption.rs:281-325.  all_cts: list of (c1, c2)."""
    if len(all_cts) == 0:
        raise PvwError("InvalidParameters", "No ciphertexts provided")
    if len(all_cts) != P.n:
        raise PvwError("InvalidParameters", f"Expected {P.n} ciphertexts, got {len(all_cts)}")
    if party_index >= P.n:
        raise PvwError("InvalidParameters", f"Party index {party_index} exceeds maximum {P.n - 1}")
    return [decrypt_party_value(P, c1, c2, sk_coeffs, party_index) for (c1, c2) in all_cts]


# --------------------------------------------------------------------------------------------------
# Seeded CRS (src/params/crs.rs:45-90).  fhe-math / rand internals recalled: PARITY UNPINNED.
# --------------------------------------------------------------------------------------------------
def chacha8_from_seed(seed32: bytes) -> "ChaCha8Rng":
    """ChaCha8Rng::from_seed"""
    return ChaCha8Rng(bytes(seed32))


def uniform_u64_sample(rng: "ChaCha8Rng", q: int) -> int:
    """rand 0.8.5 Uniform::<u64>::from(0..q).sample(rng)"""
    ints_to_reject = (MASK64 - q + 1) % q
    zone = MASK64 - ints_to_reject
    while True:
        m = rng.next_u64() * q
        if (m & MASK64) <= zone:
            return m >> 64


def poly_random_from_seed(P: "Params", seed32: bytes) -> Poly:
    """fhe-math Poly::random_from_seed(ctx, Ntt, seed): ChaCha8(SHA-256(seed)), each RNS row = Uniform(0..q_j) samples"""
    import hashlib
    prng = chacha8_from_seed(hashlib.sha256(bytes(seed32)).digest())
    return [[uniform_u64_sample(prng, q) for _ in range(P.l)] for q in P.moduli]


def crs_new_deterministic(P: "Params", seed32: bytes) -> List[List[Poly]]:
    """PvwCrs::new_deterministic, crs.rs:45-67: master ChaCha8 -> gen::<[u8; 32]>() per element (one u32 per byte)"""
    master = chacha8_from_seed(seed32)
    A = []
    for _ in range(P.k):
        row = []
        for _ in range(P.k):
            element_seed = bytes(master.next_u32() & 0xFF for _ in range(32))
            row.append(poly_random_from_seed(P, element_seed))
        A.append(row)
    return A


def siphash13_zero_key(msg: bytes) -> int:
    """Rust DefaultHasher::new() = SipHash-1-3 with k0 = k1 = 0"""
    rotl = lambda x, b: ((x << b) | (x >> (64 - b))) & MASK64
    v = [0x736F6D6570736575, 0x646F72616E646F6D, 0x6C7967656E657261, 0x7465646279746573]

    def rnd():
        v[0] = (v[0] + v[1]) & MASK64; v[1] = rotl(v[1], 13); v[1] ^= v[0]; v[0] = rotl(v[0], 32)
        v[2] = (v[2] + v[3]) & MASK64; v[3] = rotl(v[3], 16); v[3] ^= v[2]
        v[0] = (v[0] + v[3]) & MASK64; v[3] = rotl(v[3], 21); v[3] ^= v[0]
        v[2] = (v[2] + v[1]) & MASK64; v[1] = rotl(v[1], 17); v[1] ^= v[2]; v[2] = rotl(v[2], 32)

    n = len(msg)
    full = n - n % 8
    for i in range(0, full, 8):
        m = int.from_bytes(msg[i:i + 8], "little")
        v[3] ^= m; rnd(); v[0] ^= m
    last = ((n & 0xFF) << 56) | int.from_bytes(msg[full:], "little")
    v[3] ^= last; rnd(); v[0] ^= last
    v[2] ^= 0xFF
    rnd(); rnd(); rnd()
    return v[0] ^ v[1] ^ v[2] ^ v[3]


def crs_seed_from_tag(tag: str) -> bytes:
    """new_from_tag, crs.rs:74-90: hash of (tag + "CRS") as a str (bytes + 0xff), u64 cycled to 32 bytes"""
    h = siphash13_zero_key((tag + "CRS").encode() + b"\xff")
    return h.to_bytes(8, "little") * 4


# --------------------------------------------------------------------------------------------------
# Synthetic inputs, flat index = row-major position in the named array
# --------------------------------------------------------------------------------------------------
def synth_crs(P: Params, seed=DEFAULT_SEED) -> List[List[Poly]]:
    """A[i][j] = uniform residues in NTT form (like Poly::random(.., Ntt, ..), crs.rs:32); layout [k][k][L][l]."""
    k, L, l = P.k, P.L, P.l
    return [[[[uniform_residue(stream_u64(seed, TAG_A, ((i * k + j) * L + a) * l + c), P.moduli[a])
               for c in range(l)] for a in range(L)] for j in range(k)] for i in range(k)]


def synth_small(P: Params, tag: int, rows: int, cols: int, kind: str, bound=None, seed=DEFAULT_SEED, row0=0):
    """rows x cols x l small signed ints.  kind: 'cbd' (secret_variance) or 'uniform' (in [-bound, bound])."""
    l = P.l
    out = []
    for a in range(row0, row0 + rows):
        row = []
        for b in range(cols):
            vals = []
            for c in range(l):
                u = stream_u64(seed, tag, (a * cols + b) * l + c)
                vals.append(cbd(u, P.secret_variance) if kind == "cbd" else uniform_sym(u, bound))
            row.append(vals)
        out.append(row)
    return out


def synth_messages(P: Params, D: int, mode="example", seed=DEFAULT_SEED):
    """m[d][p]: 'example' = d*1000 + p + 1 (examples/pvw.rs:98-100); 'u63' = uniform below 2^63."""
    if mode == "example":
        return [[d * 1000 + p + 1 for p in range(P.n)] for d in range(D)]
    return [[stream_u64(seed, TAG_M, d * P.n + p) >> 1 for p in range(P.n)] for d in range(D)]


# --------------------------------------------------------------------------------------------------
# Named parameter sets
# --------------------------------------------------------------------------------------------------
EX_MODULI = [0xFFFFC4001, 0x1FFFFE0001]                                  # examples/pvw.rs:28-32
TEST_MODULI = [0xFFFFEE001, 0xFFFFC4001, 0x1FFFFE0001]                   # tests/crypto.rs:50-52
VD_MODULI = [0x800000022A0001, 0x800000021A0001, 0x80000002120001, 0x80000001F60001]  # examples/pvw_valid_dec.rs:40-45


def largest_ntt_primes(count: int, bits: int = 62, two_l: int = 64) -> List[int]:
    """the `count` largest primes below 2^bits that are 1 mod two_l."""
    out = []
    c = ((1 << bits) - 1) // two_l * two_l + 1
    while len(out) < count:
        if c < (1 << bits) and is_prime(c):
            out.append(c)
        c -= two_l
    return out
