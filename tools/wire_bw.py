"""Wire-format kernels (csrc/wire.cu) at the headline configuration: serialise / parse D ciphertexts of C3 (n=4096, k=256,
l=8, 17 x 62-bit) between the device store and a device byte buffer; achieved GB/s on algorithmic bytes (residues read or
written + wire bytes written or read) against the measured HBM copy bandwidth, and the host-buffer variant (PCIe inside).
usage: python tools/wire_bw.py [D] [reps]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import pvw_oracle as O  # noqa: E402
import pvw_rs_b200 as pvw  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n, k, l, L = 4096, 256, 8, 17
dev = torch.device("cuda:0")
eng = pvw.Engine(n, k, l, O.largest_ntt_primes(L))
g = torch.Generator(device=dev)
g.manual_seed(5)


def rand_polys(count):
    t = torch.empty((count, L, l), dtype=torch.int64, device=dev)
    for j, q in enumerate(eng.moduli):
        t[:, j, :] = torch.randint(0, q, (count, l), device=dev, generator=g, dtype=torch.int64)
    return t


eng.crs_upload(rand_polys(k * k).reshape(k, k, L, l))
for p0 in range(0, n, 512):
    eng.pk_upload_rows(p0, rand_polys(512 * k).reshape(512, k, L, l))          # uniform B: same traffic as genuine keys
cbd = lambda shape: (lambda b: (b & 1) - ((b >> 1) & 1))(torch.randint(0, 4, shape, device=dev, generator=g, dtype=torch.int64))
uni = lambda shape, b: torch.randint(-b, b + 1, shape, device=dev, generator=g, dtype=torch.int64)
eng.ct_reserve(D)
m = torch.randint(0, 2 ** 62, (D, n), device=dev, generator=g, dtype=torch.int64)
eng.encrypt_batch(0, m, cbd((D, k, l)), uni((D, k, l), 100), uni((D, n, l), 200))
eng.synchronize()
lay = eng.wire_layout
buf = torch.empty(D * lay.ciphertext_bytes, dtype=torch.uint8, device=dev)
c1_before, c2_before = eng.ct_download(D - 1)
res = {"D": D, "ciphertext_bytes": lay.ciphertext_bytes, "raw_bytes": (k + n) * L * l * 8, "record_bytes": lay.record_bytes}
for name, fn in (("serialize", lambda: eng.wire_ct_serialize(0, D, out=buf)), ("deserialize", lambda: eng.wire_ct_deserialize(0, D, buf))):
    for _ in range(2):
        fn()
    eng.set_option("profile", 2)
    for _ in range(reps):
        fn()
    prof = eng.profile()
    eng.set_option("profile", 0)
    ms, cnt, byts = prof["wire"]
    res[name] = {"ms_per_call": ms / reps, "launches_per_call": cnt / reps, "algorithmic_GB_per_call": byts / reps / 1e9,
                 "achieved_GBps": byts / (ms * 1e-3) / 1e9}
c1_after, c2_after = eng.ct_download(D - 1)
assert (c1_after == c1_before).all() and (c2_after == c2_before).all()
# host-buffer calls (PCIe + staging inside), pinned destination
host = torch.empty((D, lay.ciphertext_bytes), dtype=torch.uint8).pin_memory().numpy()
for name, fn in (("serialize_to_host", lambda: eng.wire_ct_serialize(0, D, out=host)), ("deserialize_from_host", lambda: eng.wire_ct_deserialize(0, D, host))):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    res[name] = {"ms_per_call": dt * 1e3, "wire_GBps": D * lay.ciphertext_bytes / dt / 1e9}
assert (host.reshape(-1) == buf.cpu().numpy()).all()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
for key in ("serialize", "deserialize"):
    res[key]["frac_of_measured_hbm"] = res[key]["achieved_GBps"] / peaks["hbm_gbs"]
print(json.dumps({"n": n, "k": k, "l": l, "L": L, "hbm_gbs": peaks["hbm_gbs"], **res}))
