"""CPU: the arithmetic identity behind the tensor-core product (pvw-rs_b200/csrc/imma.cu), modelled with numpy / Python integers.

The kernel never multiplies 62-bit numbers: it multiplies byte planes with u8 x u8 -> s32 MMAs on overlapping windows of one
accumulator, and the epilogue recombines 15 diagonal sums.  This test restates exactly that data flow -- byte planes of both
operands, window s starting DT*s columns in, B-tile row (t, d) landing in column DT*(s+t)+d, the word-grouped recombination of
`combine160` and the two-step reduction -- and checks it against plain modular dot products.  It needs no GPU; the GPU parity
tests (tests/test_gpu_imma.py) check the kernel itself."""
import numpy as np
import pytest

import pvw_oracle as O

DT = 32                      # dealers per tile
DIAGS = 15


def byte_planes(x):          # x: uint64 [..., k] -> uint8 [8, ..., k]   (plane s = byte s of every value)
    return np.stack([((x >> np.uint64(8 * s)) & np.uint64(0xFF)).astype(np.int64) for s in range(8)])


def windowed_accumulator(M, V):
    """M: uint64 [rows][k], V: uint64 [DT][k]  ->  int64 [rows][15*DT] exactly as the MMAs leave TMEM"""
    rows, k = M.shape
    Mb = byte_planes(M)                                   # [8][rows][k]
    Vb = byte_planes(V)                                   # [8][DT][k]
    B = Vb.reshape(8 * DT, k)                             # B tile: row (t, d) = t*DT + d
    acc = np.zeros((rows, 16 * DT), dtype=np.int64)       # 512 TMEM columns
    for s in range(8):                                    # one chain of MMAs per byte plane of M, on the window DT*s .. DT*s + 8*DT
        acc[:, DT * s:DT * s + 8 * DT] += Mb[s] @ B.T
    assert acc.max() < 2 ** 31                            # the s32 accumulator never overflows
    return acc[:, :DIAGS * DT]


def combine160(s):
    """the epilogue's recombination: diagonals u = r mod 4 are word aligned among themselves (imma.cu combine160)"""
    w, acc = [], 0
    get = lambda i: int(s[i]) if 0 <= i < DIAGS else 0
    fsl = lambda hi, lo, n: ((hi << n) | (lo >> (32 - n))) & 0xFFFFFFFF      # __funnelshift_l(lo, hi, n): upper word of (hi:lo) << n
    for i in range(5):
        c0, c1, c2, c3 = (get(4 * i + r) if i < 4 else 0 for r in range(4))
        p1, p2, p3 = (get(4 * i - 3), get(4 * i - 2), get(4 * i - 1)) if i > 0 else (0, 0, 0)
        if i == 4:
            p3 = 0                                                         # s[15] does not exist
        acc += c0 + fsl(c1, p1, 8) + fsl(c2, p2, 16) + fsl(c3, p3, 24)
        w.append(acc & 0xFFFFFFFF)
        acc >>= 32
    return w


@pytest.mark.parametrize("k,q", [(256, O.largest_ntt_primes(1)[0]), (250, O.EX_MODULI[1]), (18, O.TEST_MODULI[0]), (1024, O.VD_MODULI[0])])
def test_byte_plane_windows_reproduce_the_modular_product(k, q):
    rng = np.random.default_rng(k)
    rows = 5
    M = rng.integers(0, q, size=(rows, k), dtype=np.uint64)
    V = rng.integers(0, q, size=(DT, k), dtype=np.uint64)
    M[0, :] = q - 1                                       # extreme operands
    V[0, :] = q - 1
    acc = windowed_accumulator(M, V)
    for r in range(rows):
        for d in (0, 1, DT - 1):
            sums = [acc[r, DT * u + d] for u in range(DIAGS)]                # what the lane of row r reads for dealer d
            words = combine160(sums)
            value = sum(wd << (32 * i) for i, wd in enumerate(words))
            exact = sum(int(a) * int(b) for a, b in zip(M[r], V[d]))
            assert value == exact                                            # the 160-bit integer IS the unreduced dot product
            assert value % q == sum(int(a) * int(b) % q for a, b in zip(M[r], V[d])) % q
            assert words[4] < 2 ** 16                                        # the fifth word stays small: value < k * 2^124


def test_first_k_step_split_needs_no_clearing():
    """window s >= 1 touches DT columns no earlier MMA wrote: issuing its first K step as N = 7*DT accumulating plus N = DT
    overwriting equals accumulating into a cleared accumulator (the kernel's alternative to zeroing TMEM)"""
    rng = np.random.default_rng(7)
    k, rows, q = 64, 3, O.largest_ntt_primes(1)[0]
    M = rng.integers(0, q, size=(rows, k), dtype=np.uint64)
    V = rng.integers(0, q, size=(DT, k), dtype=np.uint64)
    want = windowed_accumulator(M, V)
    Mb, B = byte_planes(M), byte_planes(V).reshape(8 * DT, k)
    acc = rng.integers(-2 ** 31, 2 ** 31, size=(rows, 16 * DT)).astype(np.int64)   # garbage left by the previous tile
    for s in range(8):
        for k0 in range(0, k, 32):                                           # K = 32 bytes per MMA
            prod = Mb[s][:, k0:k0 + 32] @ B[:, k0:k0 + 32].T
            lo = DT * s
            if k0:
                acc[:, lo:lo + 8 * DT] += prod
            elif s == 0:
                acc[:, lo:lo + 8 * DT] = prod                                # the tile's very first MMA overwrites its whole window
            else:
                acc[:, lo:lo + 7 * DT] += prod[:, :7 * DT]                   # N = 7*DT, accumulate
                acc[:, lo + 7 * DT:lo + 8 * DT] = prod[:, 7 * DT:]           # N = DT (the t = 7 rows of B), overwrite
    assert (acc[:, :DIAGS * DT] == want).all()


M64 = (1 << 64) - 1


def reduce128_model(h, lo, q):
    """modarith.cuh reduce128 with 64-bit wrap-around: Barrett estimate from the two high words of mu = floor(2^128 / q)"""
    mu = (1 << 128) // q
    mu_hi, mu_lo = mu >> 64, mu & M64
    qh = (h * mu_hi + ((h * mu_lo) >> 64) + ((lo * mu_hi) >> 64)) & M64
    r = (lo - qh * q) & M64
    assert r < 4 * q                                      # what the two conditional subtractions below rely on
    if r >= 2 * q:
        r -= 2 * q
    if r >= q:
        r -= q
    return r


def reduce160_q62_model(value, q):
    """imma.cu reduce160_q62: the bits above 2^124 folded down with 2^124 mod q, then ONE Barrett step (moduli >= 2^61)"""
    w = [(value >> (32 * i)) & 0xFFFFFFFF for i in range(5)]
    c124 = (1 << 124) % q
    h = (w[3] >> 28) | ((w[4] << 4) & 0xFFFFFFFF)
    lo = (w[1] << 32) | w[0]
    hi = ((w[3] & 0x0FFFFFFF) << 32) | w[2]
    p0, p1 = h * (c124 & 0xFFFFFFFF), h * (c124 >> 32)
    a = (lo + p0) & M64
    b = (a + ((p1 << 32) & M64)) & M64
    t_hi = hi + (p1 >> 32) + (1 if a < lo else 0) + (1 if b < a else 0)
    assert t_hi < 1 << 64 and (t_hi << 64 | b) == (value & ((1 << 124) - 1)) + h * c124     # t, exactly, in two words
    assert ((t_hi << 64) | b) // q < 1 << 64              # the quotient fits one word: q >= 2^61, t < 2^125
    return reduce128_model(t_hi, b, q)


@pytest.mark.parametrize("q", [O.largest_ntt_primes(34)[0], O.largest_ntt_primes(34)[-1], (1 << 61) + 1, (1 << 62) - 57, (1 << 61) + 12345678901])
def test_one_step_reduction_of_the_160_bit_sums(q):
    """every modulus >= 2^61 (the library checks before it picks this form), sums up to 4096 * (2^64 - 1)^2 -- whatever bytes
    the operands hold -- and the extremes of every field"""
    assert q >= 1 << 61
    rng = np.random.default_rng(q % 1000003)
    top = 4096 * M64 * M64                                 # k <= 4096 terms of arbitrary 64-bit operands: < 2^140
    cases = [0, 1, q - 1, q, (1 << 124) - 1, 1 << 124, (1 << 128) - 1, 1 << 128, top, 4096 * (q - 1) ** 2, 256 * (q - 1) ** 2]
    cases += [int(rng.integers(0, 1 << 62)) * int(rng.integers(0, 1 << 62)) * int(rng.integers(1, 4097)) for _ in range(3000)]
    cases += [int.from_bytes(rng.bytes(18), "little") % (top + 1) for _ in range(3000)]
    for v in cases:
        assert reduce160_q62_model(v, q) == v % q


def test_ternary_transform_is_two_table_rows():
    """ntt.cu's table path: the transform is linear, so NTT(x) of a ternary polynomial is the sum of the rows of its four-coefficient
    groups, row index = sum (x_i + 1) 3^i over the group (hostparams.hpp builds the rows with the same host transform)"""
    P = O.Params(4, 2, 8, O.TEST_MODULI, error_bound_1=50, error_bound_2=50)
    rng = np.random.default_rng(3)
    groups = P.l // 4
    table = {}
    for g in range(groups):
        for row in range(81):
            coeffs, r = [0] * P.l, row
            for i in range(4):
                coeffs[4 * g + i] = r % 3 - 1
                r //= 3
            table[g, row] = P.ntt_forward(P.from_coefficients(coeffs))
    for _ in range(40):
        x = rng.integers(-1, 2, size=P.l)
        want = P.ntt_forward(P.from_coefficients([int(v) for v in x]))
        rows = [table[g, sum((int(x[4 * g + i]) + 1) * 3 ** i for i in range(4))] for g in range(groups)]
        for j, q in enumerate(P.moduli):
            got = [sum(int(np.asarray(r_)[j][t]) for r_ in rows) % q for t in range(P.l)]
            assert got == [int(v) for v in np.asarray(want)[j]]
