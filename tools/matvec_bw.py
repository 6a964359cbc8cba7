"""Single-call encrypt / decrypt (D = 1: the reference's own call granularity, HBM-bound matrix-vector products):
achieved GB/s of the MAC kernel on algorithmic bytes against the measured HBM copy bandwidth.
usage: python tools/matvec_bw.py [n] [reps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import pvw_oracle as O  # noqa: E402
import pvw_rs_b200 as pvw  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
k, l, L = 256, 8, 17
dev = torch.device("cuda:0")
eng = pvw.Engine(n, k, l, O.largest_ntt_primes(L))
for kv in filter(None, os.environ.get("PVW_OPTS", "").split(",")):
    a, b = kv.split("=")
    eng.set_option(a, int(b))
g = torch.Generator(device=dev)
g.manual_seed(3)
A = torch.empty((k, k, L, l), dtype=torch.int64, device=dev)
for j, q in enumerate(eng.moduli):
    A[:, :, j, :] = torch.randint(0, q, (k, k, l), device=dev, generator=g, dtype=torch.int64)
eng.crs_upload(A)
cbd = lambda shape: (lambda b: (b & 1) - ((b >> 1) & 1))(torch.randint(0, 4, shape, device=dev, generator=g, dtype=torch.int64))
uni = lambda shape, b: torch.randint(-b, b + 1, shape, device=dev, generator=g, dtype=torch.int64)
sk = cbd((n, k, l))
for p0 in range(0, n, 512):
    eng.keygen_batch(p0, sk[p0:p0 + 512].contiguous(), uni((min(512, n - p0), k, l), 100))
eng.ct_reserve(4)
m = torch.randint(0, 2 ** 62, (1, n), device=dev, generator=g, dtype=torch.int64)
r, e1, e2 = cbd((1, k, l)), uni((1, k, l), 100), uni((1, n, l), 200)
parties = np.arange(n, dtype=np.uint32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > L2: evict B between calls
for _ in range(3):
    eng.encrypt_batch(0, m, r, e1, e2)
    out = eng.decrypt_batch(parties, sk, D=1)
eng.synchronize()
assert bool((out.t() == m).all().item())
res = {}
for name, fn in (("encrypt D=1", lambda: eng.encrypt_batch(0, m, r, e1, e2)), ("decrypt D=1 (all parties)", lambda: eng.decrypt_batch(parties, sk, D=1))):
    eng.set_option("profile", 2)
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        fn()
    prof = eng.profile()
    eng.set_option("profile", 0)
    ms, cnt, byts = prof["mac_gemm"]
    res[name] = {"mac_gemm_ms_per_call": ms / reps, "launches_per_call": cnt / reps, "algorithmic_GB_per_call": byts / reps / 1e9,
                 "achieved_GBps": byts / (ms * 1e-3) / 1e9, "all_kernels_ms_per_call": sum(v[0] for v in prof.values()) / reps}
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
for v in res.values():
    v["frac_of_measured_hbm"] = v["achieved_GBps"] / peaks["hbm_gbs"]
print(json.dumps({"n": n, "k": k, "l": l, "L": L, "hbm_gbs": peaks["hbm_gbs"], **res}))
