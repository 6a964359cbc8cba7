"""Do the CUDA-core kernels (NTT, decode chain) overlap with the persistent tensor-core kernel when they run on another stream?
Two contexts on one GPU (each has its own stream): context A encrypts (c2 product on the tensor cores), context B decrypts.
usage: python tools/overlap_probe.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import pvw_oracle as O  # noqa: E402
import pvw_rs_b200 as pvw  # noqa: E402

n, k, l, L, D = 4096, 256, 8, 17, 256
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(5)
cbd = lambda shape: (lambda b: (b & 1) - ((b >> 1) & 1))(torch.randint(0, 4, shape, device=dev, generator=g, dtype=torch.int64))
uni = lambda shape, b: torch.randint(-b, b + 1, shape, device=dev, generator=g, dtype=torch.int64)


def make():
    eng = pvw.Engine(n, k, l, O.largest_ntt_primes(L))
    A = torch.empty((k, k, L, l), dtype=torch.int64, device=dev)
    for j, q in enumerate(eng.moduli):
        A[:, :, j, :] = torch.randint(0, q, (k, k, l), device=dev, generator=g, dtype=torch.int64)
    eng.crs_upload(A)
    sk = cbd((n, k, l))
    for p0 in range(0, n, 512):
        eng.keygen_batch(p0, sk[p0:p0 + 512].contiguous(), uni((512, k, l), 100))
    eng.ct_reserve(D)
    m = torch.randint(0, 2 ** 62, (D, n), device=dev, generator=g, dtype=torch.int64)
    args = (m, cbd((D, k, l)), uni((D, k, l), 100), uni((D, n, l), 200))
    eng.encrypt_batch(0, *args)
    out = torch.empty((n, D), dtype=torch.int64, device=dev)
    eng.decrypt_batch(np.arange(n, dtype=np.uint32), sk, D=D, out=out)
    eng.synchronize()
    assert bool((out.t() == m).all().item())
    return eng, args, sk, out


STAGES = int(sys.argv[1]) if len(sys.argv) > 1 else 0
a, aargs, ask, aout = make()
a.set_option("imma_stages", STAGES)
b, bargs, bsk, bout = make()
parties = np.arange(n, dtype=np.uint32)
# The Engine wrapper orders every device-tensor call after torch's current stream and torch's stream after the call, which
# serialises two contexts through that stream.  The inputs here are long finished: drop the ordering so that the two library
# streams are really independent (the first version of this probe did not, and "measured" no overlap for that reason alone).
torch.cuda.synchronize()
for e in (a, b):
    e._before_device_call = lambda: None
    e._after_device_call = lambda: None


def t(fn, reps=5):
    fn()
    a.synchronize(); b.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    a.synchronize(); b.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


enc = lambda: a.encrypt_batch(0, *aargs)
dec = lambda: b.decrypt_batch(parties, bsk, D=D, out=bout)
both = lambda: (enc(), dec())
print({"imma_stages": STAGES, "encrypt_ms": t(enc), "decrypt_ms": t(dec), "both_concurrent_ms": t(both)})
# a pure CUDA-core kernel on the other stream: the wire-format packer of context B against context A's tensor-core product
buf = torch.empty(D * b.wire_layout.ciphertext_bytes, dtype=torch.uint8, device=dev)
ser = lambda: b.wire_ct_serialize(0, D, out=buf)
print({"encrypt_ms": t(enc), "serialize_ms": t(ser), "both_concurrent_ms": t(lambda: (enc(), ser(), ser()))})
# a memory-bound torch kernel (torch's own stream, tiny shared-memory / register footprint) against the tensor-core product
x = torch.empty(1 << 28, dtype=torch.float32, device=dev)
side = torch.cuda.Stream(device=dev)
def axpy():
    with torch.cuda.stream(side):
        for _ in range(4):
            x.mul_(1.0001)
def ta(fn, reps=5):
    fn(); a.synchronize(); side.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    a.synchronize(); side.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print({"encrypt_ms": ta(enc), "torch_elementwise_ms": ta(axpy), "both_concurrent_ms": ta(lambda: (enc(), axpy()))})
