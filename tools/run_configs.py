"""Runs the BASELINE.json configurations that are not the bench headline, each with its own correctness check, and prints
one JSON line per configuration (kept under profiles/).  GPU box only.

  C1  examples/pvw.rs defaults (n=7, k=32, l=8, 2 moduli)            -- bit-exact vs the C oracle, all shares
  C2  P128, n=1024 parties, share-distribution encrypt of 1024 dealers -- first 4 dealers bit-exact vs the C oracle
  C4  P256 (k=512, l=16, 34 x 62-bit), n=8192, the row shard of rank 0 of 8 (1024 rows) on this GPU -- plaintexts recovered
  C5  pvw_valid_dec-style (k=1024, l=8, 4 x 56-bit, variance 10, bounds (1, 1172385)): every party decrypts only a random
      "valid" subset of the dealers (dealer index lists) -- subset == the matching entries of the full decryption

usage: python tools/run_configs.py [C1 C2 C4 C5]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import c_oracle as CO  # noqa: E402
import pvw_oracle as O  # noqa: E402
import pvw_rs_b200 as pvw  # noqa: E402

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev)


def cbd(shape, variance=0.5):
    if abs(variance - 0.5) < 1e-6:
        b = torch.randint(0, 4, shape, device=dev, generator=gen, dtype=torch.int64)
        return (b & 1) - ((b >> 1) & 1)
    v = int(variance)
    a = torch.randint(0, 2, shape + (2 * v,), device=dev, generator=gen, dtype=torch.int64).sum(-1)
    b = torch.randint(0, 2, shape + (2 * v,), device=dev, generator=gen, dtype=torch.int64).sum(-1)
    return a - b


def uni(shape, b):
    return torch.randint(-b, b + 1, shape, device=dev, generator=gen, dtype=torch.int64)


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, time.perf_counter() - t0


def device_system(eng, n, k, l, moduli, variance, b1, row0, nrows, seed):
    """CRS + genuine keys for rows [row0, row0+nrows) generated on the device"""
    for kv in filter(None, os.environ.get("PVW_OPTS", "").split(",")):
        name, val = kv.split("=")
        eng.set_option(name.strip(), int(val))
    gen.manual_seed(seed)
    A = torch.empty((k, k, len(moduli), l), dtype=torch.int64, device=dev)
    for j, q in enumerate(moduli):
        A[:, :, j, :] = torch.randint(0, q, (k, k, l), device=dev, generator=gen, dtype=torch.int64)
    eng.crs_upload(A)
    del A
    sk = cbd((nrows, k, l), variance)
    for p0 in range(0, nrows, 256):
        cnt = min(256, nrows - p0)
        eng.keygen_batch(row0 + p0, sk[p0:p0 + cnt].contiguous(), uni((cnt, k, l), b1))
    eng.synchronize()
    return sk


def c1():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _cases import System, engine_kwargs, params
    P = params("EX")
    S = System(P, P.n, "example")
    eng = pvw.Engine(**engine_kwargs(P))
    eng.crs_upload(S.A)
    eng.keygen_batch(0, S.sk, S.ke)
    eng.ct_reserve(P.n)
    eng.encrypt_batch(0, S.m, S.r, S.e1, S.e2)                               # warm-up: scratch buffers are allocated on first use
    eng.decrypt_batch(np.arange(P.n), S.sk, D=P.n)
    _, t_enc = timed(lambda: eng.encrypt_batch(0, S.m, S.r, S.e1, S.e2))
    got, t_dec = timed(lambda: eng.decrypt_batch(np.arange(P.n), S.sk, D=P.n))
    c1_, c2_ = S.encrypt()
    ok = all((eng.ct_download(d)[0] == c1_[d]).all() and (eng.ct_download(d)[1] == c2_[d]).all() for d in range(P.n))
    ok = ok and (got == S.co.decrypt(S.sk, c1_, c2_)).all() and (got == S.m.T).all()
    return {"config": "C1 examples/pvw.rs defaults", "n": P.n, "k": P.k, "l": P.l, "L": P.L, "bit_exact_vs_oracle": bool(ok),
            "encrypt_ms": 1e3 * t_enc, "decrypt_ms": 1e3 * t_dec}


def c2():
    n, k, l, D = 1024, 256, 8, 1024
    moduli = O.largest_ntt_primes(17)
    eng = pvw.Engine(n, k, l, moduli)
    sk = device_system(eng, n, k, l, moduli, 0.5, 100, 0, n, 11)
    eng.ct_reserve(D)
    m = torch.randint(0, 2 ** 62, (D, n), device=dev, generator=gen, dtype=torch.int64)
    r, e1, e2 = cbd((D, k, l)), uni((D, k, l), 100), uni((D, n, l), 200)
    eng.encrypt_batch(0, m, r, e1, e2)                                       # warm-up
    eng.decrypt_batch(np.arange(n, dtype=np.uint32), sk, D=D)
    _, t_enc = timed(lambda: eng.encrypt_batch(0, m, r, e1, e2))
    # first 4 dealers against the CPU oracle on identical keys / randomness / messages
    P = O.Params(n, k, l, moduli, psi=eng.psi)
    co = CO.COracle(P)
    A_h, B_h = eng.crs_download(), eng.pk_download_rows(0, n)
    h = lambda t: t[:4].cpu().numpy()
    c1_, c2_ = co.encrypt(A_h, B_h, h(m).view(np.uint64), h(r), h(e1), h(e2))
    ok = all((eng.ct_download(d)[0] == c1_[d]).all() and (eng.ct_download(d)[1] == c2_[d]).all() for d in range(4))
    out, t_dec = timed(lambda: eng.decrypt_batch(np.arange(n, dtype=np.uint32), sk, D=D))
    ok = ok and bool((out.t() == m).all().item())
    return {"config": "C2 P128 n=1024 share distribution", "n": n, "k": k, "l": l, "L": 17, "dealers": D,
            "bit_exact_vs_oracle_first_4_dealers": bool(ok), "encrypt_s": t_enc, "decrypt_s": t_dec,
            "shares_per_s_encrypt": D * n / t_enc, "shares_per_s_decrypt": D * n / t_dec}


def c4():
    n, k, l, world, D = 8192, 512, 16, 8, 64
    moduli = O.largest_ntt_primes(34)
    plan = pvw.sharding.ShardPlan(n, world, 0)
    eng = pvw.Engine(n, k, l, moduli, row0=plan.row0, nrows=plan.nrows)
    assert eng.verify_correctness_condition()
    sk = device_system(eng, n, k, l, moduli, 0.5, 100, plan.row0, plan.nrows, 12)
    eng.ct_reserve(D)
    m = torch.randint(0, 2 ** 62, (D, plan.nrows), device=dev, generator=gen, dtype=torch.int64)
    r, e1, e2 = cbd((D, k, l)), uni((D, k, l), 100), uni((D, plan.nrows, l), 200)
    pidx = np.arange(plan.row0, plan.row0 + plan.nrows, dtype=np.uint32)
    eng.encrypt_batch(0, m, r, e1, e2)
    eng.decrypt_batch(pidx, sk, D=D)                                     # warm-up: scratch buffers are allocated on first use
    _, t_enc = timed(lambda: eng.encrypt_batch(0, m, r, e1, e2))
    eng.set_option("profile", 2)
    out, t_dec = timed(lambda: eng.decrypt_batch(pidx, sk, D=D))
    prof = {k_: round(v[0], 3) for k_, v in eng.profile().items() if v[1]}
    eng.set_option("profile", 0)
    ok = bool((out.t() == m).all().item())
    return {"config": "C4 P256 n=8192, row shard of rank 0 of 8 on one GPU", "n": n, "k": k, "l": l, "L": 34, "rows": plan.nrows,
            "dealers": D, "all_shares_recovered": ok, "encrypt_s": t_enc, "decrypt_s": t_dec,
            "shares_per_s_per_gpu": D * plan.nrows / (t_enc + t_dec), "decrypt_kernels_ms": prof, "B_shard_GB": plan.nrows * k * 34 * l * 8 / 1e9}


def c5(n=1024):
    k, l, variance, b1, b2 = 1024, 8, 10.0, 1, 1172385                       # examples/pvw_valid_dec.rs:40-52
    D = min(n, 256)
    eng = pvw.Engine(n, k, l, O.VD_MODULI, secret_variance=variance, error_bound_1=b1, error_bound_2=b2)
    assert eng.verify_correctness_condition()
    sk = device_system(eng, n, k, l, O.VD_MODULI, variance, b1, 0, n, 13)
    eng.ct_reserve(D)
    d_idx = torch.arange(D, device=dev, dtype=torch.int64).reshape(D, 1)
    p_idx = torch.arange(n, device=dev, dtype=torch.int64).reshape(1, n)
    m = d_idx * 1000 + p_idx + 1                                              # examples/pvw.rs:98-100 share pattern
    r, e1, e2 = cbd((D, k, l), variance), uni((D, k, l), b1), uni((D, n, l), b2)
    mc = m.contiguous()
    parties = np.arange(n, dtype=np.uint32)
    eng.encrypt_batch(0, mc, r, e1, e2)                                      # warm-up (scratch buffers are sized on first use)
    eng.decrypt_batch(parties, sk, D=D)
    _, t_enc = timed(lambda: eng.encrypt_batch(0, mc, r, e1, e2))
    full, t_full = timed(lambda: eng.decrypt_batch(parties, sk, D=D))
    # threshold-style subset: t = ceil(2D/5) + U[0, D - t] "valid" dealers, in random order (pvw_valid_dec.rs:161-195)
    rng = np.random.default_rng(5)
    t = -(-2 * D // 5)
    valid = rng.permutation(D)[: t + rng.integers(0, D - t + 1)].astype(np.uint32)
    eng.decrypt_batch(parties, sk, dealer_slots=valid)
    sub, t_sub = timed(lambda: eng.decrypt_batch(parties, sk, dealer_slots=valid))
    ok = bool((full.t() == m).all().item()) and bool((sub == full[:, torch.from_numpy(valid.astype(np.int64)).to(dev)]).all().item())
    return {"config": f"C5 pvw_valid_dec-style subset decryption n={n}", "n": n, "k": k, "l": l, "L": 4, "dealers": D,
            "valid_dealers": int(len(valid)), "subset_equals_full_and_recovered": ok, "encrypt_s": t_enc, "decrypt_all_s": t_full,
            "decrypt_subset_s": t_sub, "shares_per_s_subset_decrypt": len(valid) * n / t_sub,
            "shares_per_s_encrypt_plus_full_decrypt": D * n / (t_enc + t_full)}


if __name__ == "__main__":
    which = sys.argv[1:] or ["C1", "C2", "C4", "C5"]
    for name in which:
        fn = {"C1": c1, "C2": c2, "C4": c4, "C5": c5, "C5_4096": lambda: c5(4096), "C5_16384": lambda: c5(16384)}[name]
        print(json.dumps(fn()), flush=True)
        torch.cuda.empty_cache()
