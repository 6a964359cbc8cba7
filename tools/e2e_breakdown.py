"""Wall-clock breakdown of the host-buffer path vs the device-resident path (where does e2e lose time?)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import pvw_oracle as O
import pvw_rs_b200 as pvw
n, k, l, L, D = 4096, 256, 8, 17, 256
dev = torch.device("cuda:0")
eng = pvw.Engine(n, k, l, O.largest_ntt_primes(L))
g = torch.Generator(device=dev); g.manual_seed(3)
A = torch.empty((k, k, L, l), dtype=torch.int64, device=dev)
for j, q in enumerate(eng.moduli):
    A[:, :, j, :] = torch.randint(0, q, (k, k, l), device=dev, generator=g, dtype=torch.int64)
eng.crs_upload(A)
cbd = lambda shape: (lambda b: (b & 1) - ((b >> 1) & 1))(torch.randint(0, 4, shape, device=dev, generator=g, dtype=torch.int64))
uni = lambda shape, b: torch.randint(-b, b + 1, shape, device=dev, generator=g, dtype=torch.int64)
sk = cbd((n, k, l))
for p0 in range(0, n, 512):
    eng.keygen_batch(p0, sk[p0:p0 + 512].contiguous(), uni((512, k, l), 100))
eng.ct_reserve(D)
m = torch.randint(0, 2 ** 62, (D, n), device=dev, generator=g, dtype=torch.int64)
r, e1, e2 = cbd((D, k, l)), uni((D, k, l), 100), uni((D, n, l), 200)
out = torch.empty((n, D), dtype=torch.int64, device=dev)
pin = lambda t: t.cpu().pin_memory()
hm, hr, he1, he2, hsk = pin(m), pin(r), pin(e1), pin(e2), pin(sk)
nm, nr, ne1, ne2, nsk = hm.numpy().view(np.uint64), hr.numpy(), he1.numpy(), he2.numpy(), hsk.numpy()
hout = torch.empty((n, D), dtype=torch.int64).pin_memory(); nout = hout.numpy().view(np.uint64)
parties = np.arange(n, dtype=np.uint32)
def wall(fn, reps=10):
    fn(); torch.cuda.synchronize(); eng.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    eng.synchronize(); torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps
print("encrypt device", wall(lambda: eng.encrypt_batch(0, m, r, e1, e2)))
print("encrypt host  ", wall(lambda: eng.encrypt_batch(0, nm, nr, ne1, ne2)))
print("decrypt device", wall(lambda: eng.decrypt_batch(parties, sk, D=D, out=out)))
print("decrypt host  ", wall(lambda: eng.decrypt_batch(parties, nsk, D=D, out=nout)))
buf = torch.empty(64 << 20, dtype=torch.uint8, device=dev); hb = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
def h2d():
    buf.copy_(hb, non_blocking=True)
print("H2D 64 MiB pinned GB/s", 64 * 1.048576e-3 / (wall(h2d) * 1e-3) / 1e0 * 1e-0)
def d2h():
    hb.copy_(buf, non_blocking=True)
print("D2H 64 MiB pinned GB/s", 64 * 1.048576e-3 / (wall(d2h) * 1e-3))
for name, fn in (("encrypt device", lambda: eng.encrypt_batch(0, m, r, e1, e2)), ("encrypt host", lambda: eng.encrypt_batch(0, nm, nr, ne1, ne2)),
                 ("decrypt device", lambda: eng.decrypt_batch(parties, sk, D=D, out=out)), ("decrypt host", lambda: eng.decrypt_batch(parties, nsk, D=D, out=nout))):
    eng.set_option("profile", 2)
    w = wall(fn, 5)
    pr = eng.profile()
    eng.set_option("profile", 0)
    print(name, "wall", round(w, 3), "kernels", {k_: round(v[0] / 6, 3) for k_, v in pr.items() if v[1]}, "sum", round(sum(v[0] for v in pr.values()) / 6, 3))
