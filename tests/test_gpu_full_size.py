"""GPU, BASELINE.json's full sizes (C3: 128-bit set, n = 4096 parties): size-independent properties where the oracle is
too slow to check every value -- encrypt -> decrypt round trip of every share, additivity of the ciphertexts in
(m, r, e1, e2), agreement of a row shard with the unsharded key, and a sampled bit-exact comparison with the oracle."""
import numpy as np
import pytest

import c_oracle as CO
import pvw_oracle as O

pytestmark = pytest.mark.gpu

N, K, ELL, LIMBS, D = 4096, 256, 8, 17, 8


@pytest.fixture(scope="module")
def system():
    import torch
    import pvw_rs_b200 as pvw
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    moduli = O.largest_ntt_primes(LIMBS)
    eng = pvw.Engine(N, K, ELL, moduli)
    A = torch.empty((K, K, LIMBS, ELL), dtype=torch.int64, device=dev)
    for j, q in enumerate(moduli):
        A[:, :, j, :] = torch.randint(0, q, (K, K, ELL), device=dev, generator=g, dtype=torch.int64)
    eng.crs_upload(A)
    cbd = lambda shape: (lambda b: (b & 1) - ((b >> 1) & 1))(torch.randint(0, 4, shape, device=dev, generator=g, dtype=torch.int64))
    uni = lambda shape, b: torch.randint(-b, b + 1, shape, device=dev, generator=g, dtype=torch.int64)
    sk, ke = cbd((N, K, ELL)), uni((N, K, ELL), 100)
    for p0 in range(0, N, 512):
        eng.keygen_batch(p0, sk[p0:p0 + 512].contiguous(), ke[p0:p0 + 512].contiguous())
    eng.ct_reserve(3 * D)
    inputs = []
    for _ in range(2):
        inputs.append((torch.randint(0, 2 ** 61, (D, N), device=dev, generator=g, dtype=torch.int64), cbd((D, K, ELL)),
                       uni((D, K, ELL), 100), uni((D, N, ELL), 200)))
    return pvw, torch, eng, A, sk, ke, inputs, moduli


def test_round_trip_every_share(system):
    pvw, torch, eng, A, sk, ke, inputs, moduli = system
    m, r, e1, e2 = inputs[0]
    eng.encrypt_batch(0, m, r, e1, e2)
    out = eng.decrypt_batch(np.arange(N, dtype=np.uint32), sk, D=D)
    eng.synchronize()
    assert bool((out.t() == m).all().item())
    # a permuted subset of parties and dealers gives the matching entries (examples/pvw_valid_dec.rs:198-210)
    ps = np.array([4095, 0, 2048, 17, 1023], dtype=np.uint32)
    ds = np.array([7, 0, 3], dtype=np.uint32)
    sub = eng.decrypt_batch(ps, sk[torch.from_numpy(ps.astype(np.int64)).to(sk.device)].contiguous(), dealer_slots=ds)
    eng.synchronize()
    assert bool((sub.cpu() == m.cpu()[ds.astype(np.int64)][:, ps.astype(np.int64)].t()).all().item())


def test_ciphertexts_are_additive(system):
    """Enc(m1; r1, e1, e2) + Enc(m2; r2, e1', e2') == Enc(m1 + m2; r1 + r2, e1 + e1', e2 + e2') slot-wise mod q_j."""
    pvw, torch, eng, A, sk, ke, inputs, moduli = system
    (m1, r1, a1, b1), (m2, r2, a2, b2) = inputs
    eng.encrypt_batch(0, m1, r1, a1, b1)
    eng.encrypt_batch(D, m2, r2, a2, b2)
    eng.encrypt_batch(2 * D, m1 + m2, r1 + r2, a1 + a2, b1 + b2)
    q = np.array(moduli, dtype=np.uint64).reshape(1, LIMBS, 1)
    for d in (0, D - 1):
        x1, y1 = eng.ct_download(d)
        x2, y2 = eng.ct_download(D + d)
        x3, y3 = eng.ct_download(2 * D + d)
        assert ((x1 + x2) % q == x3).all() and ((y1 + y2) % q == y3).all()      # residues < 2^62: no u64 overflow


def test_row_shard_agrees_with_full_key(system):
    pvw, torch, eng, A, sk, ke, inputs, moduli = system
    m, r, e1, e2 = inputs[0]
    eng.encrypt_batch(0, m, r, e1, e2)
    plan = pvw.sharding.ShardPlan(N, 8, 5)
    lo, hi = plan.row0, plan.row0 + plan.nrows
    shard = pvw.Engine(N, K, ELL, moduli, row0=lo, nrows=hi - lo)
    shard.crs_upload(A)
    shard.keygen_batch(lo, sk[lo:hi].contiguous(), ke[lo:hi].contiguous())
    shard.ct_reserve(D)
    shard.encrypt_batch(0, m[:, lo:hi].contiguous(), r, e1, e2[:, lo:hi].contiguous(), c1_range=plan.dealer_slice(D))
    dlo, dhi = plan.dealer_slice(D)
    for d in range(D):
        full_c1, full_c2 = eng.ct_download(d)
        c1, c2 = shard.ct_download(d)
        assert (c2 == full_c2[lo:hi]).all()
        if dlo <= d < dhi:
            assert (c1 == full_c1).all()
        else:
            shard.ct_upload(d, c1=full_c1)                       # what the all-gather delivers
    out = shard.decrypt_batch(np.arange(lo, hi, dtype=np.uint32), sk[lo:hi].contiguous(), D=D)
    shard.synchronize()
    assert bool((out.t() == m[:, lo:hi]).all().item())


def test_sampled_values_match_the_oracle(system):
    """bit-exact c1 (all rows) and 64 sampled rows of c2 / plaintexts of one dealer against the C restatement"""
    pvw, torch, eng, A, sk, ke, inputs, moduli = system
    m, r, e1, e2 = inputs[1]
    eng.encrypt_batch(0, m, r, e1, e2)
    rows = np.sort(np.random.default_rng(9).choice(N, 64, replace=False))
    P = O.Params(64, K, ELL, moduli, psi=eng.psi)                 # the oracle sees the 64 sampled parties as its n
    co = CO.COracle(P)
    A_h = eng.crs_download()
    B_h = np.concatenate([eng.pk_download_rows(int(p), 1) for p in rows])
    h = lambda t: t.cpu().numpy()
    d = 3
    c1_o, c2_o = co.encrypt(A_h, B_h, h(m)[d:d + 1, rows].view(np.uint64), h(r)[d:d + 1], h(e1)[d:d + 1], h(e2)[d:d + 1, rows])
    c1_g, c2_g = eng.ct_download(d)
    assert (c1_g == c1_o[0]).all() and (c2_g[rows] == c2_o[0]).all()
    dec = co.decrypt(h(sk)[rows], c1_o, c2_o)
    assert (dec[:, 0] == h(m)[d, rows].view(np.uint64)).all()
