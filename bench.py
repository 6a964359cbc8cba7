#!/usr/bin/env python
"""bench.py -- party-shares encrypted + decrypted per second (BASELINE.json), B200 arm and CPU reference arm.

A "step" is one pass of the hot path over one batch of synthetic dealers:
    encrypt (c1 = A r + e1, c2 = B r + e2 + m g) for the step's dealers x all local parties, then
    decrypt (<s, c1> - c2, l-redundant decode) of every local party x the step's dealers (C5: x a random "valid" subset).
Workloads (BASELINE.json `configs`), selected with --config; C3 is the one the metric is quoted on:
    C1  examples/pvw.rs defaults             n=7,    k=32,   l=8,  2 moduli
    C2  128-bit set                          n=1024, k=256,  l=8,  17 x 62-bit
    C3  128-bit set (headline, default)      n=4096            (--parties 8192: the north_star target size)
    C4  256-bit set                          n=8192, k=512,  l=16, 34 x 62-bit, rows of B sharded over the GPUs
    C5  examples/pvw_valid_dec.rs:40-52      k=1024, l=8, 4 x 56-bit, variance 10, bounds (1, 1 172 385); n=4096 by default,
        --parties 1024 / 16384 for the sweep; every party decrypts only a random valid subset of the dealers (:161-210)
Multi-GPU (one process per GPU, torchrun): rows of B / parties are sharded across ranks, A is broadcast once over NCCL,
every rank computes c1 for its slice of the step's dealers and the slices are exchanged -- by default with copy-engine peer
copies ordered by stream counters (pvw_shard_*, under the c2 product), optionally with an NCCL all-gather or by replicating
the c1 product.  Dealers per step grow with the rank count so that the per-GPU work is fixed ("weak").

  python bench.py [--config C3] [--gpus N] [--steps K] [--warmup W] [--dealers D_per_gpu] [--parties n] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = the same through the host-buffer C-ABI
calls (H2D of m / r / e1 / e2 / sk and D2H of the plaintexts inside the timed region), `roofline` = the dominant kernel
against the tensor-pipe / HBM peak, `cpu_baseline` = the oracle port on the host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from dataclasses import dataclass
from typing import List

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "party-shares encrypted+decrypted/sec"
UNIT = "shares/s"
SEED = 0x5056572D42323030


@dataclass
class Workload:
    key: str
    title: str
    n: int
    k: int
    l: int
    moduli: List[int]
    variance: float
    b1: int
    b2: int
    dealers: int          # per GPU per step
    msg: str              # "u62": uniform 62-bit messages; "share": d*1000 + p + 1 (examples/pvw.rs:98-100)
    subset: bool          # decrypt a random valid subset of the dealers (examples/pvw_valid_dec.rs:161-210)
    cpu_dealers: int      # dealers of the cpu_baseline sample
    cpu_rows: int         # parties of the cpu_baseline sample (extrapolated linearly to n when smaller)

    @property
    def L(self):
        return len(self.moduli)

    def name(self, n):
        qbits = sum(q.bit_length() for q in self.moduli)
        what = "encrypt + subset decrypt (valid dealers only)" if self.subset else "encrypt + all-party decrypt"
        return f"{self.title}: n={n} parties, k={self.k}, l={self.l}, L={self.L} moduli (Q ~{qbits} bit); {what}"

    def bytes_per_share(self, n):
        """algorithmic bytes per encrypted + decrypted share."""
        poly = 8 * self.L * self.l
        enc = ((self.k * self.k + n * self.k) * poly + (self.k + n) * poly + 8 * (2 * self.k * self.l + n * self.l + n)) / n
        dec = self.k * poly + poly + 8
        return enc + dec


def workloads():
    import pvw_oracle as O          # parameter-set definitions only (prime search), not on any timed path
    p128 = O.largest_ntt_primes(17)
    return {
        "C1": Workload("C1", "C1 examples/pvw.rs defaults", 7, 32, 8, list(O.EX_MODULI), 0.5, 50, 50, 7, "share", False, 7, 7),
        "C2": Workload("C2", "C2 P128", 1024, 256, 8, p128, 0.5, 100, 200, 256, "u62", False, 128, 1024),
        "C3": Workload("C3", "C3 P128", 4096, 256, 8, p128, 0.5, 100, 200, 256, "u62", False, 64, 4096),
        "C4": Workload("C4", "C4 P256", 8192, 512, 16, O.largest_ntt_primes(34), 0.5, 100, 200, 128, "u62", False, 8, 4096),
        "C5": Workload("C5", "C5 pvw_valid_dec-style", 4096, 1024, 8, list(O.VD_MODULI), 10.0, 1, 1172385, 256, "share", True, 64, 4096),
    }


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port (the reference itself is Rust + un-vendored fhe-math: not buildable here).  One sample =
# c1 for Dc dealers, c2 + decrypt for Dc dealers x n_s parties; a step of the real workload costs
#     t_c1 + (n / n_s) * (t_c2 + frac_dec * t_dec)        (every term is linear in the parties; frac_dec = valid / all dealers)
# ----------------------------------------------------------------------------------------------------------------
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample(co, A, B, sk, m, r, e1, e2, n_total, frac_dec=1.0):
    """times the three legs on B [n_s][k][L][l], sk [n_s][k][l], m [Dc][n_s], e2 [Dc][n_s][l]; returns (shares/s, seconds, c1, c2, dec)"""
    Dc, n_s = m.shape
    t0 = time.perf_counter()
    c1, _ = co.encrypt(A, B[:1], m[:, :1], r, e1, e2[:, :1], want_c2=False)
    t1 = time.perf_counter()
    _, c2 = co.encrypt(A, B, m, r, e1, e2, want_c1=False)
    t2 = time.perf_counter()
    dec = co.decrypt(sk, c1, c2)
    t3 = time.perf_counter()
    step_s = (t1 - t0) + (n_total / n_s) * ((t2 - t1) + frac_dec * (t3 - t2))
    return Dc * n_total / step_s, t3 - t0, c1, c2, dec


def cpu_uniform_inputs(P, W, Dc, n_s):
    import c_oracle as CO
    import pvw_oracle as O
    A = CO.synth_crs_np(P)
    # uniform rows: identical arithmetic and memory traffic, plaintexts are not recovered
    u = CO.stream_np(SEED, O.TAG_B, 0, n_s * P.k * P.L * P.l).reshape(n_s, P.k, P.L, P.l)
    B = CO._mulhi_np(u, np.broadcast_to(np.array(P.moduli, dtype=np.uint64).reshape(1, 1, P.L, 1), u.shape))
    kind = "cbd"
    sk = CO.synth_small_np(P, O.TAG_SK, n_s, P.k, kind)
    m = CO.synth_messages_np(P, Dc, "u63")[:, :n_s]
    r = CO.synth_small_np(P, O.TAG_R, Dc, P.k, kind)
    e1 = CO.synth_small_np(P, O.TAG_E1, Dc, P.k, "uniform", W.b1)
    e2 = CO.synth_small_np(P, O.TAG_E2, Dc, n_s, "uniform", W.b2)
    return A, B, sk, np.ascontiguousarray(m), r, e1, e2


def run_reference(args, W: Workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import c_oracle as CO
    import pvw_oracle as O
    n = args.parties or W.n
    n_s = min(n, W.cpu_rows)
    P = O.Params(n_s, W.k, W.l, W.moduli, secret_variance=W.variance, error_bound_1=W.b1, error_bound_2=W.b2)
    co = CO.COracle(P)
    cores = host_cores()
    co.set_threads(cores)          # torch.distributed.run exports OMP_NUM_THREADS=1: the reference's rayon pool uses every core
    D = args.cpu_dealers or max(1, W.cpu_dealers // 4)
    frac = subset_fraction(W, W.dealers) if W.subset else 1.0
    A, B, sk, m, r, e1, e2 = cpu_uniform_inputs(P, W, D, n_s)
    for _ in range(args.warmup):
        cpu_sample(co, A, B[:8], sk[:8], m[:1, :8], r[:1], e1[:1], e2[:1, :8], n, frac)
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        vals.append(cpu_sample(co, A, B, sk, m, r, e1, e2, n, frac)[0])
    dt = time.perf_counter() - t0
    val = len(vals) / sum(1.0 / v for v in vals)                    # shares / total extrapolated step time
    sample = (f"per step: c1 for {D} dealers + (c2, decrypt) for {D} dealers x {n_s} of the {n} parties, scaled linearly to n "
              f"({'valid-subset fraction %.3f of the decryptions; ' % frac if W.subset else ''}uniform synthetic B); {co.threads} OpenMP threads; "
              f"{dt:.1f} s for {args.steps} steps")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * D * n / val, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": W.name(n), "dealers_per_step": D, "dealers_per_step_note": "the metric is per share: the B200 arm batches "
                       f"{W.dealers} dealers per GPU and step of the same workload, this arm a bounded sample of it",
                       "note": "CPU port of the reference path (oracle/pvw_oracle.c, OpenMP over dealers x parties like the crate's rayon loops); "
                               "the Rust crate and its fhe-math dependency cannot be built in this image"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": co.threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def subset_fraction(W: Workload, D: int) -> float:
    return len(valid_subset(D)) / D


def valid_subset(D: int) -> np.ndarray:
    """threshold-style subset of a step's dealers (pvw_valid_dec.rs:161-195): t + U[0, D - t] of them, t = ceil(2 D / 5), random order"""
    rng = np.random.default_rng(5)
    t = -(-2 * D // 5)
    return rng.permutation(D)[: t + rng.integers(0, D - t + 1)].astype(np.uint32)


# ----------------------------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------------------------
def synth_device(torch, dev, gen, shape, kind, bound=None, variance=0.5):
    if kind == "cbd":
        if abs(variance - 0.5) < 1e-6:   # CBD(0.5): {-1, 0, 1}
            b = torch.randint(0, 4, shape, device=dev, generator=gen, dtype=torch.int64)
            return (b & 1) - ((b >> 1) & 1)
        v = int(variance)                 # CBD(v): popcount(2v bits) - popcount(2v bits)  (uniform.rs:38-67)
        a = torch.randint(0, 2, tuple(shape) + (2 * v,), device=dev, generator=gen, dtype=torch.int8).sum(-1, dtype=torch.int64)
        b = torch.randint(0, 2, tuple(shape) + (2 * v,), device=dev, generator=gen, dtype=torch.int8).sum(-1, dtype=torch.int64)
        return a - b
    if kind == "uniform":
        return torch.randint(-bound, bound + 1, shape, device=dev, generator=gen, dtype=torch.int64)
    if kind == "u62":
        return torch.randint(0, 2 ** 62, shape, device=dev, generator=gen, dtype=torch.int64)
    raise ValueError(kind)


def run_b200(args, W: Workload):
    import torch
    import torch.distributed as dist
    import pvw_rs_b200 as pvw

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    n, k, l, L = args.parties or W.n, W.k, W.l, W.L
    moduli = W.moduli
    plan = pvw.sharding.ShardPlan(n, world, rank)
    nrows, row0 = plan.nrows, plan.row0
    Dg = args.dealers or W.dealers    # dealers per GPU per step (this rank's c1 slice)
    D = Dg * world                    # dealers per step
    c1_lo, c1_hi = plan.dealer_slice(D)
    exchange = args.c1_exchange if world > 1 else "none"

    eng = pvw.Engine(n, k, l, moduli, secret_variance=W.variance, error_bound_1=W.b1, error_bound_2=W.b2, row0=row0, nrows=nrows, device=local)
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    for kv in filter(None, os.environ.get("PVW_OPTS", "").split(",")):     # tuning knobs, e.g. PVW_OPTS=imma=0
        name, val = kv.split("=")
        eng.set_option(name.strip(), int(val))

    # ---- setup (untimed): CRS broadcast over NCCL, genuine public keys generated on the device ----------------
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED & 0x7FFFFFFF)
    A = torch.empty((k, k, L, l), dtype=torch.int64, device=dev)
    if rank == 0:
        for j, q in enumerate(moduli):
            A[:, :, j, :] = torch.randint(0, q, (k, k, l), device=dev, generator=gen, dtype=torch.int64)
    if world > 1:
        dist.broadcast(A, 0)
    torch.cuda.synchronize()
    eng.crs_upload(A)
    del A
    gen.manual_seed(1000 + rank)
    sk = synth_device(torch, dev, gen, (nrows, k, l), "cbd", variance=W.variance)
    kchunk = max(1, min(512, (256 << 20) // (k * L * l * 8)))
    for p0 in range(0, nrows, kchunk):
        cnt = min(kchunk, nrows - p0)
        ke = synth_device(torch, dev, gen, (cnt, k, l), "uniform", W.b1)
        eng.keygen_batch(row0 + p0, sk[p0:p0 + cnt].contiguous(), ke)
    eng.synchronize()
    eng.ct_reserve(D)
    xch = pvw.sharding.CopyEngineExchange(eng, plan, device=dev) if exchange == "ce" else None

    # ---- the step's synthetic inputs: same dealers on every rank (r, e1), local columns of m / e2 ---------------
    # the dealers' randomness is sampled once (rank 0) and broadcast over NCCL / NVLink: every rank encrypts under the SAME r, e1
    # (north_star: "A and r are replicated via NCCL broadcast"); m and e2 are per party, i.e. local to the rank that owns the rows
    gen.manual_seed(77 + 1000 * rank)
    r = synth_device(torch, dev, gen, (D, k, l), "cbd", variance=W.variance)
    e1 = synth_device(torch, dev, gen, (D, k, l), "uniform", W.b1)
    if world > 1:
        dist.broadcast(r, 0)
        dist.broadcast(e1, 0)
    gen.manual_seed(78 + rank)
    if W.msg == "share":
        m = (torch.arange(D, device=dev, dtype=torch.int64).reshape(D, 1) * 1000 + torch.arange(row0, row0 + nrows, device=dev, dtype=torch.int64).reshape(1, nrows) + 1).contiguous()
    else:
        m = synth_device(torch, dev, gen, (D, nrows), "u62")
    e2 = synth_device(torch, dev, gen, (D, nrows, l), "uniform", W.b2)
    valid = valid_subset(D) if W.subset else None
    Dv = len(valid) if W.subset else D
    out = torch.empty((nrows, Dv), dtype=torch.int64, device=dev)
    parties = np.arange(row0, row0 + nrows, dtype=np.uint32)
    c1_view = eng.c1_store_tensor(0, D) if exchange == "nccl" else None
    vt = torch.from_numpy(valid.astype(np.int64)).to(dev) if W.subset else None

    def step(m_, r_, e1_, e2_, sk_, out_):
        """encrypt + c1 exchange + decrypt; the arguments are all device tensors or all host arrays"""
        if exchange == "ce":        # one call: c1 slice, its peer copies queued on the copy engines (PVW_ENC_PUSH_C1), then the c2 product over them
            eng.encrypt_batch(0, m_, r_, e1_, e2_, c1_range=(c1_lo, c1_hi), push_c1=True)
            xch.wait()
        elif exchange == "nccl":    # one call, then the in-place all-gather ordered on the library's stream
            eng.encrypt_batch(0, m_, r_, e1_, e2_, c1_range=(c1_lo, c1_hi))
            with torch.cuda.stream(ext):
                pvw.sharding.all_gather_c1(c1_view, plan)
        elif exchange == "replicate":   # every rank computes the whole of c1: no exchange, world x the c1 product
            eng.encrypt_batch(0, m_, r_, e1_, e2_)
        else:
            eng.encrypt_batch(0, m_, r_, e1_, e2_, c1_range=(c1_lo, c1_hi))
        res = eng.decrypt_batch(parties, sk_, D=None if W.subset else D, dealer_slots=valid, out=out_)
        if exchange == "ce":
            xch.release()
        return res

    def step_device():
        return step(m, r, e1, e2, sk, out)

    # host-buffer path (e2e): pinned host inputs, H2D inside the library calls, plaintexts read back to the host.
    # Narrow element types (PVW_IN_SECRET_I8 / PVW_IN_ERROR_*): the values are tiny, the reference's i64 is 8x / 2-4x the bytes.
    etype = torch.int16 if max(W.b1, W.b2) < (1 << 15) else torch.int32
    pin = lambda t: t.cpu().pin_memory()
    h64 = {"m": pin(m), "r": pin(r), "e1": pin(e1), "e2": pin(e2), "sk": pin(sk)}
    hn = {"m": h64["m"], "r": pin(r.to(torch.int8)), "e1": pin(e1.to(etype)), "e2": pin(e2.to(etype)), "sk": pin(sk.to(torch.int8))}
    n64 = {kk: v.numpy() for kk, v in h64.items()}
    nn = {kk: v.numpy() for kk, v in hn.items()}
    n_m = n64["m"].view(np.uint64)
    h_out = torch.empty((nrows, Dv), dtype=torch.int64).pin_memory()
    n_out = h_out.numpy().view(np.uint64)

    def step_host():
        return step(n_m, nn["r"], nn["e1"], nn["e2"], nn["sk"], n_out)

    def step_host_i64():
        return step(n_m, n64["r"], n64["e1"], n64["e2"], n64["sk"], n_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(ext)
        for _ in range(steps):
            fn()
        b.record(ext)
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- correctness guard: the plaintexts of the step are the messages (genuine keys) -------------------------
    want = m if not W.subset else m[vt]
    step_device()
    eng.synchronize()
    bad_dev = (out.t() != want)
    res_host = step_host().copy()
    bad_host = torch.from_numpy(res_host.view(np.int64).T != want.cpu().numpy())
    for name, bad in (("device-resident", bad_dev), ("host-buffer", bad_host)):
        if bool(bad.any().item()):
            idx = bad.nonzero()
            raise SystemExit(f"bench: rank {rank}: {idx.shape[0]} decrypted shares differ from the messages on the {name} path "
                             f"(first (dealer, party) = {idx[0].tolist()}, dealers hit: {idx[:, 0].unique().numel()}, parties hit: "
                             f"{idx[:, 1].unique().numel()}) -- refusing to report a number")

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------------
    warm = max(args.warmup, 3)
    for _ in range(warm - 1):
        step_device()
    eng.set_option("profile", 2)
    l0 = eng.launch_count
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total = timed(step_device, args.steps)
    clk = clocks.stop() if rank == 0 else None
    launches = eng.launch_count - l0
    prof = eng.profile()
    eng.set_option("profile", 0)
    shares_per_step = D * n
    value = shares_per_step * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers through the C-ABI calls ------------------------------------------------------------------
    for _ in range(2):
        step_host()
    ms_e2e = timed(step_host, args.steps)
    e2e_value = shares_per_step * args.steps / (ms_e2e * 1e-3)
    small = lambda d: d["r"].nbytes + d["sk"].nbytes + d["e2"].nbytes + d["e1"][c1_lo:c1_hi].nbytes
    h2d = n_m.nbytes + small(nn) + parties.nbytes + (valid.nbytes if W.subset else 0)
    d2h = nrows * Dv * 8
    # the same with the reference's i64 element type for every small input
    e2e_steps = max(2, args.steps // 2)
    step_host_i64()
    ms_e2e64 = timed(step_host_i64, e2e_steps)
    # ... and with every ciphertext of the step returned to the host in the crate's wire format (a caller of the reference API
    # owns PvwCiphertext values; `e2e` keeps them in the device store, where the decryption of the same box reads them)
    wire = None
    try:
        wl = eng.wire_layout
        blob = torch.empty((D, int(wl.ciphertext_bytes)), dtype=torch.uint8).pin_memory()
        nblob = blob.numpy()

        def step_wire():
            r_ = step_host()
            eng.wire_ct_serialize(0, D, out=nblob)
            return r_
        step_wire()
        ws = max(2, args.steps // 3)
        ms_wire = timed(step_wire, ws)
        wire = {"value": shares_per_step * ws / (ms_wire * 1e-3), "unit": UNIT, "ms_per_step": ms_wire / ws,
                "d2h_ciphertext_bytes_per_step": int(D * wl.ciphertext_bytes),
                "what": "e2e + bincode(PvwCiphertext) of every dealer of the step written to pinned host memory (pvw_wire_ct_serialize)"}
        del blob, nblob
    except Exception as ex:           # the wire leg is informative; never lose the headline to it
        wire = {"error": str(ex)[:200]}

    # ---- the reference's own call granularity: ONE encrypt / ONE decrypt_party_shares-per-party pass (D = 1).  The MAC kernel
    # is then a matrix-vector product that must stream B (or the secret keys) from HBM once: the HBM-bound case of the path.
    single = None
    if rank == 0 and world == 1:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
        single = {}
        out1 = torch.empty((nrows, 1), dtype=torch.int64, device=dev)
        for name, fn in (("encrypt", lambda: eng.encrypt_batch(0, m[:1], r[:1], e1[:1], e2[:1], c1_range=(0, 1))),
                         ("decrypt_all_parties", lambda: eng.decrypt_batch(parties, sk, D=1, out=out1))):
            fn()
            eng.set_option("profile", 2)
            reps = 10
            for _ in range(reps):
                flush.zero_()
                torch.cuda.synchronize()
                fn()
            pr = eng.profile()
            eng.set_option("profile", 0)
            ms1, n1, b1 = pr["mac_gemm"]
            single[name] = {"mac_gemm_ms": ms1 / reps, "algorithmic_GB": b1 / reps / 1e9, "achieved_GBps": b1 / (ms1 * 1e-3) / 1e9,
                            "all_kernels_ms": sum(v[0] for v in pr.values()) / reps}
        del flush

    if rank != 0:
        if xch is not None:
            eng.synchronize()
            barrier()
            xch.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (mac_gemm) ------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    imma_on = os.environ.get("PVW_OPTS", "").replace(" ", "").find("imma=0") < 0
    mac_ms, mac_n, mac_bytes = (prof["imma_gemm"] if imma_on and prof["imma_gemm"][1] else prof["mac_gemm"])   # tensor-core / CUDA-core product
    achieved = mac_bytes / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0
    kernel_ms = {kname: round(v[0] / args.steps, 4) for kname, v in prof.items()}

    def json_lines(path, key):
        try:
            for ln in open(os.path.join(ROOT, "profiles", path)):
                if key in ln:
                    return float(json.loads(ln)[key])
        except Exception:
            pass
        return None

    def traffic_of(path):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", path))).get("dram_bytes_per_launch")
        except Exception:
            return None
    traffic = traffic_of("r02_imma_gemm_traffic.json") or traffic_of("r01_imma_gemm_traffic.json") if W.key == "C3" else None
    int_peak = json_lines("r01_int_peaks.json", "mac_karatsuba_per_s")
    shoup_peak = json_lines("r01_int_peaks.json", "mulmod_shoup_per_s")
    mac_rate = (mac_bytes / ((k + 1.0) * 8)) * k / (mac_ms * 1e-3) if mac_ms > 0 else 0.0
    # NTT kernels against the measured Shoup modular-multiply rate.  `ntt_small` = the transforms that run butterflies whatever the
    # input: e1 (c1 finisher), e2 + m * g_hat (c2 finisher).  `ntt_planes` = secrets / randomness -> byte planes: ternary polynomials
    # (the reference's default variance) go through table lookups there, so that kind is reported with its time and the
    # butterfly-EQUIVALENT rate only, not held against the multiply peak.
    ntt_ms, ntt_n, _ = prof["ntt_small"]
    planes_ms, planes_n, _ = prof.get("ntt_planes", (0.0, 0, 0.0))
    c1_dealers = D if exchange == "replicate" else (c1_hi - c1_lo)
    log2l = l.bit_length() - 1
    bf = L * ((l // 2) * log2l)                                                          # butterflies per polynomial
    ntt_mulmods = args.steps * ((c1_dealers * k + D * nrows) * bf + D * nrows * L * l)   # e1 slice, e2, m * g_hat
    ntt_rate = ntt_mulmods / (ntt_ms * 1e-3) if ntt_ms > 0 else 0.0
    planes_equiv = args.steps * (D * k + nrows * k) * bf / (planes_ms * 1e-3) if planes_ms > 0 else 0.0   # r, sk
    # The batched product runs on the INT8 tensor cores (csrc/imma.cu): 64 u8 x u8 multiply-accumulates per 62-bit one, so the
    # kernel's roof is the tensor pipe.  Peak: the rate measured on this part with the kernel's own MMA stream and nothing else
    # (tools/csrc/imma_probe.cu mode 4: operands resident in shared memory, no epilogue -> profiles/r02_int8_peak.json) when that
    # file exists, else twice the measured cuBLAS bf16 burst (kind::i8 issues K = 32 per instruction where bf16 issues 16).
    imma_on = imma_on and prof["imma_gemm"][1] > 0
    bf16 = float(peaks.get("bf16_tflops", 2250.0))
    int8_meas = json_lines("r02_int8_peak.json", "int8_tops_mma_only")
    int8_peak = int8_meas if int8_meas else 2.0 * bf16
    int8_peak_sustained = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 2250.0)))
    int8_ops = mac_rate * 64 * 2 / 1e12
    hbm_view = {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "what": "SURVEY 8d algorithmic bytes (one operand row read per (dealer, row) + one polynomial written) over the kernel "
                        "time: > 1 because dealers are batched -- every staged tile serves 32 dealers from shared memory"}
    if imma_on:
        roofline = {"bound": "tensor", "kernel": "imma_gemm (imma_gemm_kernel: tcgen05.mma kind::i8)", "achieved": int8_ops, "peak": int8_peak,
                    "unit": "TOP/s (u8 x u8 -> s32, dense)", "frac": int8_ops / int8_peak,
                    "peak_source": ("measured on B200: the kernel's MMA stream alone, operands resident, no epilogue (profiles/r02_int8_peak.json)" if int8_meas
                                    else "2 x measured cuBLAS bf16 burst (MEASURED_PEAKS.json bf16_tflops)" if peaks else "2 x nominal dense bf16 (fallback)"),
                    "frac_of_2x_burst_bf16": int8_ops / (2.0 * bf16), "frac_of_2x_sustained_bf16": int8_ops / int8_peak_sustained,
                    "frac_of_nominal_4500_TOPs": int8_ops / 4500.0,
                    "traffic": traffic, "ops_per_launch": mac_rate * 128 * (mac_ms * 1e-3) / max(mac_n, 1),
                    "int8_macs_per_62bit_mac": 64}
    else:
        roofline = {"bound": "hbm", "kernel": "mac_gemm (IMAD kernel)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback", "traffic": traffic}
    roofline.update({
                "launches_per_step": mac_n / args.steps, "avg_launch_ms": mac_ms / max(mac_n, 1),
                "algorithmic_bytes_per_launch": mac_bytes / max(mac_n, 1), "hbm_algorithmic": hbm_view,
                "kernel_ms_per_step": kernel_ms, "kernel_share_of_step": round(mac_ms / ms_total, 4),
                "ntt": {"kernel": "ntt_small", "achieved": ntt_rate, "peak": shoup_peak, "unit": "Shoup modular multiplies/s",
                        "frac": (ntt_rate / shoup_peak) if shoup_peak else None,
                        "what": "forward butterflies + gadget multiplies of e1 and e2 (the c1 / c2 finishers) per step over those kernels' event "
                                "time; peak = register-resident mulmod_shoup loop (profiles/r01_int_peaks.json)",
                        "planes": {"kernel": "ntt_planes (r, sk -> byte planes)", "ms_per_step": round(planes_ms / args.steps, 4),
                                   "butterfly_equivalent_per_s": planes_equiv,
                                   "what": "table lookups for ternary polynomials (secret variance 0.5), butterflies otherwise; 64-bit "
                                           "secrets narrowed to one byte first: the equivalent rate is not a multiply rate"}},
                "single_call": {"what": "D = 1 (one reference-style encrypt call / one all-party decrypt pass): HBM-bound matrix-vector "
                                        "form on the CUDA cores (mac.cu), L2 flushed between calls, rows = %d" % nrows,
                                **{kname: dict(v, frac=v["achieved_GBps"] / peak) for kname, v in (single or {}).items()}},
                "modmuladds_per_s": mac_rate,
                "integer_pipe": {"achieved": mac_rate, "peak": int_peak, "unit": "62-bit modular multiply-accumulates/s",
                                 "frac": (mac_rate / int_peak) if int_peak else None,
                                 "peak_source": "ceiling of the CUDA-core form (3 IMAD.WIDE + carries per MAC, tools/csrc/int_peaks.cu, "
                                                "profiles/r01_int_peaks.json): the tensor-core kernel is measured against it for scale"},
                "note": "batched encrypt / decrypt multiply on the INT8 tensor cores: operands are byte planes, the 64 byte products of a "
                        "62-bit multiply accumulate on overlapping windows of the TMEM accumulator into 15 diagonal sums, recombined and "
                        "reduced exactly in the epilogue (DESIGN.md 4); PVW_OPTS=imma=0 runs the CUDA-core kernel instead"})

    # ---- CPU baseline: oracle port on the host cores, bounded sample, checked bit for bit against the GPU ---------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import c_oracle as CO
        import pvw_oracle as O
        n_s = min(n, W.cpu_rows)
        Dc = min(args.cpu_baseline_dealers or W.cpu_dealers, Dg)
        P = O.Params(n_s, k, l, moduli, psi=eng.psi, secret_variance=W.variance, error_bound_1=W.b1, error_bound_2=W.b2)
        co = CO.COracle(P)
        co.set_threads(host_cores())
        A_h = eng.crs_download()
        B_h = eng.pk_download_rows(0, n_s)
        frac = Dv / D
        val_c, secs, c1_h, c2_h, dec_h = cpu_sample(co, A_h, B_h, n64["sk"][:n_s], np.ascontiguousarray(n_m[:Dc, :n_s]), n64["r"][:Dc], n64["e1"][:Dc],
                                                    np.ascontiguousarray(n64["e2"][:Dc, :n_s]), n, frac)
        # the CPU port and the GPU agree on this sample (same keys, same randomness): ciphertext of dealer 0, plaintexts of the sample
        g1, g2 = eng.ct_download(0)
        if W.subset:
            cols = [i for i, d in enumerate(valid) if d < Dc]
            same_pt = bool((dec_h[:, [int(valid[i]) for i in cols]] == res_host[:n_s, cols]).all())
        else:
            same_pt = bool((dec_h == res_host[:n_s, :Dc]).all())
        same = bool(same_pt and (g1 == c1_h[0]).all() and (g2[:n_s] == c2_h[0]).all())
        cpu = {"value": val_c, "unit": UNIT, "cores": co.threads, "kind": "port",
               "sample": f"c1 for {Dc} dealers + (c2, decrypt) for {Dc} dealers x {n_s} of the {n} parties of the same workload"
                         f"{', scaled linearly to n' if n_s < n else ''}, {secs:.1f} s; bit-identical to the GPU result: {same}"}
        if not same:
            raise SystemExit("bench: CPU port and GPU disagree on the sample")

    par = {"none": "single GPU",
           "ce": f"B rows / parties sharded x{world}; c1 dealer slices exchanged by copy-engine peer copies + stream counters (pvw_shard_*) under the c2 product",
           "nccl": f"B rows / parties sharded x{world}; c1 dealer slices all-gathered (NCCL)",
           "replicate": f"B rows / parties sharded x{world}; c1 computed by every rank (no exchange)"}[exchange]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": W.name(n), "config_key": W.key, "dealers_per_step": D, "dealers_per_gpu_per_step": Dg, "shares_per_step": shares_per_step,
                       "decrypted_shares_per_step": Dv * n, "parallelism": par,
                       "l2": "inputs larger than L2 (B shard %.0f MB, ciphertext store %.0f MB per step vs 126 MB L2)" % (
                           nrows * k * L * l * 8 / 1e6, D * (nrows + k) * L * l * 8 / 1e6),
                       "keys": "genuine (device keygen); every decrypted share checked == message before timing"},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "input_types": f"m u64; r, sk int8; e1, e2 {str(etype).replace('torch.', '')} (PVW_IN_SECRET_I8 / PVW_IN_ERROR_*); ciphertexts stay in the device store",
                    "int64_inputs": {"value": shares_per_step * e2e_steps / (ms_e2e64 * 1e-3), "ms_per_step": ms_e2e64 / e2e_steps,
                                     "h2d_bytes_per_step": int(n_m.nbytes + small(n64) + parties.nbytes)},
                    "ciphertexts_to_host": wire},
            "roofline": roofline, "cpu_baseline": cpu,
            "hbm_roof_shares_per_s": peak * 1e9 / W.bytes_per_share(n) * world}
    emit(line)
    if xch is not None:
        eng.synchronize()
        barrier()
        xch.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on stdout.  Route fd 1 to
    stderr for the whole run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"], help="BASELINE.json configuration (C3 = the headline)")
    ap.add_argument("--dealers", type=int, default=0, help="dealers per GPU per step (default: the configuration's)")
    ap.add_argument("--parties", type=int, default=0, help="override n (C3: 8192 = north_star target; C5 sweep: 1024 / 4096 / 16384)")
    ap.add_argument("--c1-exchange", default="ce", choices=["ce", "nccl", "replicate"],
                    help="multi-GPU: copy-engine peer copies (default), NCCL all-gather, or replicated c1")
    ap.add_argument("--cpu-dealers", type=int, default=0, help="dealers per step of the --impl reference arm (default: a quarter of the cpu_baseline sample)")
    ap.add_argument("--cpu-baseline-dealers", type=int, default=0, help="dealers in the cpu_baseline sample (default: the configuration's, 10-20 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    W = workloads()[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "b200" and world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args, W)
    else:
        run_b200(args, W)


if __name__ == "__main__":
    main()
