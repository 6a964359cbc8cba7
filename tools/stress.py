"""Stress: repeat encrypt + decrypt on a mid-size P128 system and count iterations whose plaintexts differ from the
messages (race hunting).  Usage: python tools/stress.py [iters] [opts like gemm_tile=1,refill_lag=2]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import pvw_rs_b200 as pvw  # noqa: E402
import pvw_oracle as O  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
opts = sys.argv[2] if len(sys.argv) > 2 else ""
n, k, l, D = 1024, 256, 8, 128
dev = torch.device("cuda:0")
eng = pvw.Engine(n, k, l, O.largest_ntt_primes(17))
for kv in filter(None, opts.split(",")):
    a, b = kv.split("=")
    eng.set_option(a, int(b))
g = torch.Generator(device=dev)
g.manual_seed(1)
A = torch.empty((k, k, 17, l), dtype=torch.int64, device=dev)
for j, q in enumerate(eng.moduli):
    A[:, :, j, :] = torch.randint(0, q, (k, k, l), device=dev, generator=g, dtype=torch.int64)
eng.crs_upload(A)
cbd = lambda shape: (lambda b: (b & 1) - ((b >> 1) & 1))(torch.randint(0, 4, shape, device=dev, generator=g, dtype=torch.int64))
uni = lambda shape, b: torch.randint(-b, b + 1, shape, device=dev, generator=g, dtype=torch.int64)
sk = cbd((n, k, l))
eng.keygen_batch(0, sk, uni((n, k, l), 100))
eng.ct_reserve(D)
bad_enc = bad_dec = 0
ref_c2 = None
for it in range(iters):
    m = torch.randint(0, 2 ** 62, (D, n), device=dev, generator=g, dtype=torch.int64)
    r, e1, e2 = cbd((D, k, l)), uni((D, k, l), 100), uni((D, n, l), 200)
    eng.encrypt_batch(0, m, r, e1, e2)
    out = eng.decrypt_batch(np.arange(n, dtype=np.uint32), sk, D=D)
    eng.synchronize()
    ok = bool((out.t() == m).all().item())
    if not ok:
        # which stage? re-run the decrypt alone: if it now matches, the decrypt pass was the faulty one
        out2 = eng.decrypt_batch(np.arange(n, dtype=np.uint32), sk, D=D)
        eng.synchronize()
        if bool((out2.t() == m).all().item()):
            bad_dec += 1
        else:
            bad_enc += 1
        wrong = (out.t() != m).nonzero()
        print(f"iter {it}: {wrong.shape[0]} wrong shares, first at (dealer, party) = {wrong[0].tolist()}; decrypt-retry ok: {bool((out2.t() == m).all().item())}")
print(f"opts={opts!r} iters={iters} bad_encrypt={bad_enc} bad_decrypt={bad_dec}")
