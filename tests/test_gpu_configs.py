"""GPU: the BASELINE.json configurations round 1 left without an oracle comparison at their real inner dimensions.

  * C5 -- examples/pvw_valid_dec.rs:40-52 at its FULL k = 1024 (4 x 56-bit moduli, l = 8, variance 10, bounds (1, 1 172 385)):
    n = 130 parties (two 128-row tiles of the tensor-core product, the second ragged), D = 33 dealers (two 32-dealer tiles),
    eight 128-byte K-chunks per byte plane: the chunk-major loop of csrc/imma.cu.
  * C4 -- the 256-bit set (README.md:77-81: k = 512, l = 16, 34 x 62-bit moduli) with 130 rows of B and 33 dealers.
  * the 128-bit set at n = 8192 parties (north_star target), sampled like tests/test_gpu_full_size.py.
  * ring degrees above 32 (parameters.rs:140-144 accepts every power of two >= 8): l = 64 on the generic kernels.
Every case: keygen rows, c1, c2 and plaintexts bit-exact against the C restatement, on the tensor-core product AND on the
CUDA-core kernel, plus the subset / threshold form of examples/pvw_valid_dec.rs:161-210."""
import numpy as np
import pytest

import c_oracle as CO
import pvw_oracle as O
from _cases import P128_MODULI, P256_MODULI, System, engine_kwargs

pytestmark = pytest.mark.gpu

FULL = {
    "VD1024": lambda: O.Params(130, 1024, 8, O.VD_MODULI, secret_variance=10.0, error_bound_1=1, error_bound_2=1172385),
    "P256": lambda: O.Params(130, 512, 16, P256_MODULI),
    "P128": lambda: O.Params(130, 256, 8, P128_MODULI),
}


@pytest.fixture(scope="module")
def pkg():
    import pvw_rs_b200
    return pvw_rs_b200


@pytest.fixture(scope="module", params=list(FULL))
def case(request):
    P = FULL[request.param]()
    S = System(P, 33, "u63" if request.param != "VD1024" else "example")
    c1, c2 = S.encrypt()
    want = S.co.decrypt(S.sk, c1, c2)
    assert (want == S.m.T).all()                          # genuine keys: the oracle recovers every message
    return request.param, P, S, c1, c2, want


@pytest.mark.parametrize("imma", [1, 0])
def test_full_inner_dimension_matches_oracle(pkg, case, imma):
    name, P, S, c1, c2, want = case
    D = S.D
    eng = pkg.Engine(**engine_kwargs(P))
    eng.set_option("imma", imma)
    eng.crs_upload(S.A)
    eng.keygen_batch(0, S.sk, S.ke)                        # b = A^T s + e on the device (crs.rs:138-171)
    assert (eng.pk_download_rows(0, P.n) == S.B).all(), "keygen rows differ from the oracle"
    eng.ct_reserve(D)
    eng.set_option("profile", 2)
    eng.encrypt_batch(0, S.m, S.r, S.e1, S.e2)
    for d in range(D):
        g1, g2 = eng.ct_download(d)
        assert (g1 == c1[d]).all(), f"c1 of dealer {d}"
        assert (g2 == c2[d]).all(), f"c2 of dealer {d}"
    got = eng.decrypt_batch(np.arange(P.n), S.sk, D=D)
    assert (got == want).all()
    used_tensor_cores = eng.profile()["imma_gemm"][1] >= 1
    assert used_tensor_cores == bool(imma)
    eng.set_option("profile", 0)
    # threshold-style subset (pvw_valid_dec.rs:161-210): a random "valid" subset of the dealers, in random order
    rng = np.random.default_rng(5)
    t = -(-2 * D // 5)
    valid = rng.permutation(D)[: t + rng.integers(0, D - t + 1)].astype(np.uint32)
    sub = eng.decrypt_batch(np.arange(P.n), S.sk, dealer_slots=valid)
    assert (sub == want[:, valid]).all()
    # one recipient, every dealer (decrypt_party_shares, decryption.rs:281-325): the matrix-vector form
    one = eng.decrypt_batch(np.array([129], dtype=np.uint32), S.sk[129:130], D=D)
    assert (one == want[129:130]).all()


def test_narrow_inputs_equal_int64_inputs(pkg, case):
    """PVW_IN_SECRET_I8 / PVW_IN_ERROR_I32 / _I16: same ciphertexts and plaintexts as the reference's i64 inputs"""
    import torch
    name, P, S, c1, c2, want = case
    D = 9
    etype = np.int32 if P.error_bound_2 >= (1 << 15) else np.int16
    eng = pkg.Engine(**engine_kwargs(P))
    eng.crs_upload(S.A)
    eng.keygen_batch(0, S.sk.astype(np.int8), S.ke.astype(etype))
    assert (eng.pk_download_rows(0, P.n) == S.B).all()
    eng.ct_reserve(D)
    for device in (False, True):
        conv = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()) if device else (lambda a: a)
        m = S.m[:D]
        eng.encrypt_batch(0, conv(m.view(np.int64)) if device else m, conv(S.r[:D].astype(np.int8)), conv(S.e1[:D].astype(etype)),
                          conv(S.e2[:D].astype(etype)))
        for d in (0, D - 1):
            g1, g2 = eng.ct_download(d)
            assert (g1 == c1[d]).all() and (g2 == c2[d]).all()
        got = eng.decrypt_batch(np.arange(P.n), conv(S.sk.astype(np.int8)), D=D)
        got = got.cpu().numpy().view(np.uint64) if device else got
        assert (got == want[:, :D]).all()


def test_p128_at_8192_parties_sampled(pkg):
    """north_star target size for the 128-bit set: every share round-trips, 64 sampled rows bit-exact against the oracle"""
    import torch
    N, K, ELL, LIMBS, D = 8192, 256, 8, 17, 16
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(4242)
    eng = pkg.Engine(N, K, ELL, P128_MODULI)
    A = torch.empty((K, K, LIMBS, ELL), dtype=torch.int64, device=dev)
    for j, q in enumerate(P128_MODULI):
        A[:, :, j, :] = torch.randint(0, q, (K, K, ELL), device=dev, generator=g, dtype=torch.int64)
    eng.crs_upload(A)
    cbd = lambda shape: (lambda b: (b & 1) - ((b >> 1) & 1))(torch.randint(0, 4, shape, device=dev, generator=g, dtype=torch.int64))
    uni = lambda shape, b: torch.randint(-b, b + 1, shape, device=dev, generator=g, dtype=torch.int64)
    sk, ke = cbd((N, K, ELL)), uni((N, K, ELL), 100)
    for p0 in range(0, N, 1024):
        eng.keygen_batch(p0, sk[p0:p0 + 1024].contiguous(), ke[p0:p0 + 1024].contiguous())
    eng.ct_reserve(D)
    m = torch.randint(0, 2 ** 62, (D, N), device=dev, generator=g, dtype=torch.int64)
    r, e1, e2 = cbd((D, K, ELL)), uni((D, K, ELL), 100), uni((D, N, ELL), 200)
    eng.encrypt_batch(0, m, r, e1, e2)
    out = eng.decrypt_batch(np.arange(N, dtype=np.uint32), sk, D=D)
    eng.synchronize()
    assert bool((out.t() == m).all().item())
    rows = np.sort(np.random.default_rng(11).choice(N, 64, replace=False))
    P = O.Params(64, K, ELL, P128_MODULI, psi=eng.psi)
    co = CO.COracle(P)
    h = lambda t: t.cpu().numpy()
    A_h = eng.crs_download()
    B_h = np.concatenate([eng.pk_download_rows(int(p), 1) for p in rows])
    assert (B_h == co.keygen(A_h, h(sk)[rows], h(ke)[rows])).all()
    for d in (0, D - 1):
        c1_o, c2_o = co.encrypt(A_h, B_h, h(m)[d:d + 1, rows].view(np.uint64), h(r)[d:d + 1], h(e1)[d:d + 1], h(e2)[d:d + 1, rows])
        c1_g, c2_g = eng.ct_download(d)
        assert (c1_g == c1_o[0]).all() and (c2_g[rows] == c2_o[0]).all()
        assert (co.decrypt(h(sk)[rows], c1_o, c2_o)[:, 0] == h(out)[rows, d].view(np.uint64)).all()


@pytest.mark.parametrize("ell,imma", [(64, 1), (64, 0), (128, 0), (256, 1)])
def test_ring_degree_above_32(pkg, ell, imma):
    P = O.Params(9, 6, ell, O.largest_ntt_primes(3 if ell < 256 else 5, 62, 2 * ell), error_bound_1=50, error_bound_2=50)
    D = 10
    S = System(P, D, "u63")
    c1, c2 = S.encrypt()
    want = S.co.decrypt(S.sk, c1, c2)
    eng = pkg.Engine(**engine_kwargs(P))
    eng.set_option("imma", imma)
    eng.set_option("imma_min_rows", 1)
    eng.crs_upload(S.A)
    eng.keygen_batch(0, S.sk, S.ke)
    assert (eng.pk_download_rows(0, P.n) == S.B).all()
    eng.ct_reserve(D)
    eng.encrypt_batch(0, S.m, S.r, S.e1, S.e2)
    for d in range(D):
        g1, g2 = eng.ct_download(d)
        assert (g1 == c1[d]).all() and (g2 == c2[d]).all()
    assert (eng.decrypt_batch(np.arange(P.n), S.sk, D=D) == want).all()
    assert (eng.ntt_forward_small(S.sk[0]) == np.stack([S.co.ntt_small(S.sk[0, j]) for j in range(P.k)])).all()
