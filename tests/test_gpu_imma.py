"""GPU: the tensor-core form of the matrix product (csrc/imma.cu: tcgen05.mma kind::i8 on byte diagonals) against the oracle
and against the IMAD kernel, bit-exact, through the C ABI.  The default build uses it whenever a call batches >= 8 dealers;
here it is forced on for every batch size, with 16-dealer chunks so that the chunk loops run too."""
import numpy as np
import pytest

import pvw_oracle as O
from _cases import P128_MODULI, P256_MODULI, System, engine_kwargs, params

pytestmark = pytest.mark.gpu

# name -> oracle parameters: the shared sets with an even k, plus ragged ones (nothing divides a tile) for this kernel
SETS = {
    "EX": lambda: params("EX"),
    "T16": lambda: params("T16"),
    "VDs": lambda: params("VDs"),
    "P128s": lambda: params("P128s"),
    "P256s": lambda: params("P256s"),
    "RAG2": lambda: O.Params(13, 6, 8, O.TEST_MODULI, error_bound_1=50, error_bound_2=50),
    "L32b": lambda: O.Params(5, 4, 32, O.largest_ntt_primes(5), secret_variance=1.0),
    "WIDE": lambda: O.Params(300, 18, 8, O.TEST_MODULI, error_bound_1=50, error_bound_2=50),   # two row tiles, k*8 = 144 bytes
}


@pytest.fixture(scope="module")
def pkg():
    import pvw_rs_b200
    return pvw_rs_b200


def forced(pkg, P, on=True, **over):
    eng = pkg.Engine(**engine_kwargs(P, **over))
    eng.set_option("imma", 1 if on else 0)
    eng.set_option("imma_min_dealers", 1)
    eng.set_option("imma_min_rows", 1)
    eng.set_option("imma_chunk_dealers", 16)
    return eng


def load(eng, S, cap):
    eng.crs_upload(S.A)
    eng.pk_upload_rows(eng.row0, S.B[eng.row0:eng.row0 + eng.nrows])
    eng.ct_reserve(cap)
    return eng


@pytest.mark.parametrize("name,D", [("EX", 1), ("EX", 7), ("EX", 20), ("T16", 5), ("VDs", 3), ("RAG2", 19), ("L32b", 2), ("WIDE", 33),
                                    ("P128s", 5), ("P128s", 33), ("P256s", 3)])
def test_encrypt_decrypt_match_oracle_and_imad(pkg, name, D):
    P = SETS[name]()
    S = System(P, D, "u63")
    c1, c2 = S.encrypt()
    want = S.co.decrypt(S.sk, c1, c2)
    eng = load(forced(pkg, P), S, D + 1)
    launches0 = eng.launch_count
    eng.encrypt_batch(1, S.m, S.r, S.e1, S.e2)
    for d in range(D):
        g1, g2 = eng.ct_download(1 + d)
        assert (g1 == c1[d]).all(), f"c1 dealer {d}"
        assert (g2 == c2[d]).all(), f"c2 dealer {d}"
    slots = np.arange(1, D + 1, dtype=np.uint32)
    got = eng.decrypt_batch(np.arange(P.n), S.sk, dealer_slots=slots)
    assert (got == want).all()
    if name != "L32b":
        assert (got == S.m.T).all()
    # the tensor-core product ran
    eng.set_option("profile", 2)
    eng.decrypt_batch(np.arange(P.n), S.sk, dealer_slots=slots)
    assert eng.profile()["imma_gemm"][1] >= 1
    eng.set_option("profile", 0)
    # permuted subsets of dealers and parties (examples/pvw_valid_dec.rs:198-210), identical on the IMAD kernel
    ds = np.array(sorted(set([D, 1, 1 + D // 2])), dtype=np.uint32)[::-1].copy()
    ps = np.array([P.n - 1, 0, P.n // 2], dtype=np.uint32)
    sub = eng.decrypt_batch(ps, S.sk[ps], dealer_slots=ds)
    assert (sub == want[np.ix_(ps, ds - 1)]).all()
    ref = load(forced(pkg, P, on=False), S, D + 1)
    ref.encrypt_batch(1, S.m, S.r, S.e1, S.e2)
    assert (ref.decrypt_batch(ps, S.sk[ps], dealer_slots=ds) == sub).all()
    assert launches0 < eng.launch_count


def test_device_inputs_row_shards_and_dealer_slices(pkg):
    """row-sharded contexts with c1 computed for a dealer slice only, CUDA-tensor inputs"""
    import torch
    P = SETS["WIDE"]()
    D = 21
    S = System(P, D, "u63")
    c1, c2 = S.encrypt()
    want = S.co.decrypt(S.sk, c1, c2)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
    for lo, hi, dlo, dhi in ((0, 130, 0, 9), (130, 300, 9, 21)):
        eng = load(forced(pkg, P, row0=lo, nrows=hi - lo), S, D)
        if lo == 0:
            eng.encrypt_batch(0, dev(S.m[:, lo:hi]), dev(S.r), dev(S.e1), dev(S.e2[:, lo:hi]), c1_range=(dlo, dhi))
        else:   # the two halves separately (PVW_ENC_C1_ONLY / PVW_ENC_C2_ONLY), as the multi-GPU step issues them
            eng.encrypt_batch(0, None, dev(S.r), dev(S.e1), None, c1_range=(dlo, dhi), part="c1")
            eng.encrypt_batch(0, dev(S.m[:, lo:hi]), dev(S.r), None, dev(S.e2[:, lo:hi]), part="c2")
        for d in range(D):
            g1, g2 = eng.ct_download(d)
            assert (g2 == c2[d, lo:hi]).all()
            if dlo <= d < dhi:
                assert (g1 == c1[d]).all()
            else:
                eng.ct_upload(d, c1=c1[d])
        out = eng.decrypt_batch(np.arange(lo, hi), dev(S.sk[lo:hi]), D=D)
        assert (out.cpu().numpy().view(np.uint64) == want[lo:hi]).all()


@pytest.mark.parametrize("name,D", [("EX", 20), ("WIDE", 33), ("P128s", 33), ("L32b", 2)])
def test_two_sm_form_matches(pkg, name, D):
    """option imma_pair: the same product with tcgen05.mma.cta_group::2 on CTA pairs (measured alternative, off by default)"""
    P = SETS[name]()
    S = System(P, D, "u63")
    c1, c2 = S.encrypt()
    eng = load(forced(pkg, P), S, D)
    eng.set_option("imma_pair", 1)
    eng.encrypt_batch(0, S.m, S.r, S.e1, S.e2)
    for d in range(D):
        g1, g2 = eng.ct_download(d)
        assert (g1 == c1[d]).all() and (g2 == c2[d]).all(), f"dealer {d}"
    assert (eng.decrypt_batch(np.arange(P.n), S.sk, D=D) == S.co.decrypt(S.sk, c1, c2)).all()


def test_few_rows_take_the_cuda_core_kernel_by_default(pkg):
    """one party decrypting many ciphertexts (the reference's decrypt_party_shares): rows = 1 -> no tensor-core tile is worth it"""
    P = SETS["P128s"]()
    D = 40
    S = System(P, D, "u63")
    c1, c2 = S.encrypt()
    eng = load(pkg.Engine(**engine_kwargs(P)), S, D)      # default options
    eng.encrypt_batch(0, S.m, S.r, S.e1, S.e2)            # 24 rows x 40 dealers: tensor cores
    eng.set_option("profile", 2)
    one = eng.decrypt_batch(np.array([5], dtype=np.uint32), S.sk[5:6], D=D)
    assert eng.profile()["imma_gemm"][1] == 0             # the CUDA-core kernel ran
    many = eng.decrypt_batch(np.arange(P.n, dtype=np.uint32), S.sk, D=D)
    assert eng.profile()["imma_gemm"][1] >= 1             # 24 parties: tensor cores
    eng.set_option("profile", 0)
    want = S.co.decrypt(S.sk, c1, c2)
    assert (one == want[5:6]).all() and (many == want).all()


def test_odd_k_falls_back_to_the_imad_kernel(pkg):
    P = params("RAG")                                    # k = 5: rows of 40 bytes cannot be TMA sources
    S = System(P, 9, "u63")
    c1, c2 = S.encrypt()
    eng = load(forced(pkg, P), S, 9)
    eng.encrypt_batch(0, S.m, S.r, S.e1, S.e2)
    assert all((eng.ct_download(d)[0] == c1[d]).all() and (eng.ct_download(d)[1] == c2[d]).all() for d in range(9))
    assert (eng.decrypt_batch(np.arange(P.n), S.sk, D=9) == S.m.T).all()


def test_planes_only_keeps_one_copy_of_the_public_key(pkg):
    """option planes_only: the u64 operand copy of B is freed once its byte planes exist and rebuilt from them on demand"""
    P = SETS["WIDE"]()
    D = 20
    S = System(P, D, "u63")
    c1, c2 = S.encrypt()
    want = S.co.decrypt(S.sk, c1, c2)
    eng = load(forced(pkg, P), S, D)
    eng.set_option("planes_only", 1)
    eng.encrypt_batch(0, S.m, S.r, S.e1, S.e2)                       # tensor-core path: builds the planes, releases the u64 copy
    assert all((eng.ct_download(d)[1] == c2[d]).all() for d in (0, D - 1))
    assert (eng.pk_download_rows(0, P.n) == S.B).all()               # rebuilt from the planes, bit for bit
    eng.set_option("imma", 0)
    eng.encrypt_batch(0, S.m[:1], S.r[:1], S.e1[:1], S.e2[:1])       # a single call on the CUDA cores needs the operand form again
    assert (eng.ct_download(0)[1] == c2[0]).all()
    eng.set_option("imma", 1)
    eng.pk_upload_rows(3, S.B[5:6])                                  # a key update invalidates the planes; they are rebuilt from the new B
    eng.encrypt_batch(0, S.m, S.r, S.e1, S.e2)
    B2 = S.B.copy()
    B2[3] = S.B[5]
    c1b, c2b = S.co.encrypt(S.A, B2, S.m, S.r, S.e1, S.e2)
    assert all((eng.ct_download(d)[1] == c2b[d]).all() for d in (0, D - 1))
    eng.pk_upload_rows(3, S.B[3:4])
    eng.encrypt_batch(0, S.m, S.r, S.e1, S.e2)
    assert (eng.decrypt_batch(np.arange(P.n), S.sk, D=D) == want).all()


@pytest.mark.parametrize("name,stype,big", [("EX", np.int8, 0), ("EX", np.int64, 0), ("P128s", np.int8, 0), ("P128s", np.int64, 0), ("P128s", np.int64, 1),
                                            ("T16", np.int8, 0), ("T16", np.int64, 0), ("P256s", np.int8, 0), ("P256s", np.int64, 0), ("P256s", np.int64, 1)])
def test_ternary_tables_and_butterflies_agree(pkg, name, stype, big):
    """Ring degrees 8 and 16: polynomials with coefficients in {-1, 0, 1} are transformed by table lookup (ntt.cu), any other polynomial by
    the butterflies -- mixed in one call here, for every input element size (64-bit secrets are narrowed to one byte on the device when they
    fit), against the oracle and against the option switched off."""
    P = SETS[name]()
    D = 9
    S = System(P, D, "u63")
    rng = np.random.default_rng(77)
    sk = S.sk.copy()
    sk[::2, ::3, 5] = 2                                     # every third polynomial of every other party leaves the ternary set
    sk[1, 1, :8] = [-128, 127, -2, 2, 3, -3, 0, 1]          # the edges of the one-byte range
    sk[3 % P.n] = rng.integers(-1, 2, size=sk[0].shape)     # a fully ternary key that is not the seeded one
    if big:
        sk[2 % P.n, 0, 0] = 1000                            # 64-bit secrets that do not fit one byte: the narrowed copy is dropped (ntt.cu)
    r = S.r.copy()
    r[0, 0, :8] = [1, -1, 0, 2, -2, 1, 0, -1]
    r[1, 1, P.l - 1] = -2                                   # the last coefficient of the last group
    c1, c2 = S.co.encrypt(S.A, S.B, S.m, r, S.e1, S.e2)
    want = S.co.decrypt(sk, c1, c2)                         # (not the messages: sk no longer matches B -- garbage decodes, exactly)
    outs = []
    for tables in (1, 0):
        eng = load(forced(pkg, P), S, D)
        eng.set_option("ternary_tables", tables)
        etype = np.int32 if P.error_bound_2 >= (1 << 15) else np.int16
        eng.encrypt_batch(0, S.m, r.astype(stype), S.e1.astype(etype), S.e2.astype(etype))
        for d in (0, D - 1):
            g1, g2 = eng.ct_download(d)
            assert (g1 == c1[d]).all() and (g2 == c2[d]).all()
        got = eng.decrypt_batch(np.arange(P.n), sk.astype(stype), D=D)
        assert (got == want).all()
        outs.append(got)
    assert (outs[0] == outs[1]).all()
