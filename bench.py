#!/usr/bin/env python
"""bench.py -- party-shares encrypted + decrypted per second on the 128-bit parameter set (BASELINE.json).

A "step" is one pass of the hot path over one batch of synthetic dealers:
    encrypt (c1 = A r + e1, c2 = B r + e2 + m g) for the step's dealers x all local parties, then
    decrypt (<s, c1> - c2, l-redundant decode) of every local party x every dealer of the step.
Workload C3 of SURVEY.md 8(d): k=256, l=8, 17 x 62-bit moduli (Q 1054 bit), n=4096 parties.
Multi-GPU (one process per GPU, torchrun): rows of B / parties are sharded across ranks, A is broadcast once over
NCCL, every rank computes c1 for its slice of the step's dealers and the slices are all-gathered over NCCL; the
number of dealers per step grows with the rank count so that the per-GPU work is fixed ("weak").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dealers D_per_gpu] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = the same through the host-buffer C-ABI
calls (H2D of r/e1/e2/m/sk and D2H of the plaintexts inside the timed region), `roofline` = the MAC kernel against
the measured HBM peak on algorithmic bytes, `cpu_baseline` = the oracle port on the host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "party-shares encrypted+decrypted/sec"
UNIT = "shares/s"
N_PARTIES, K_DIM, ELL, N_LIMBS = 4096, 256, 8, 17
SEED = 0x5056572D42323030


def p128_moduli():
    import pvw_oracle as O          # parameter-set definition only (prime search), not on any timed path
    return O.largest_ntt_primes(N_LIMBS)


def workload_name(n=N_PARTIES):
    return f"C3 P128: n={n} parties, k={K_DIM}, l={ELL}, L={N_LIMBS}x62-bit (Q 1054 bit); encrypt + all-party decrypt"


def bytes_per_share(n=N_PARTIES):
    """algorithmic bytes per encrypted + decrypted share."""
    poly = 8 * N_LIMBS * ELL
    enc = ((K_DIM * K_DIM + n * K_DIM) * poly + (K_DIM + n) * poly + 8 * (2 * K_DIM * ELL + n * ELL + n)) / n
    dec = K_DIM * poly + poly + 8
    return enc + dec


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
# reference arm / CPU baseline: the oracle port (the reference itself is Rust + un-vendored fhe-math: not buildable here)
# ----------------------------------------------------------------------------------------------------------------
def cpu_inputs(P, D, B=None):
    import c_oracle as CO
    import pvw_oracle as O
    A = CO.synth_crs_np(P)
    if B is None:   # uniform rows: identical arithmetic and memory traffic, plaintexts are not recovered
        u = CO.stream_np(SEED, O.TAG_B, 0, P.n * P.k * P.L * P.l).reshape(P.n, P.k, P.L, P.l)
        B = CO._mulhi_np(u, np.broadcast_to(np.array(P.moduli, dtype=np.uint64).reshape(1, 1, P.L, 1), u.shape))
    sk = CO.synth_small_np(P, O.TAG_SK, P.n, P.k, "cbd")
    m = CO.synth_messages_np(P, D, "u63")
    r = CO.synth_small_np(P, O.TAG_R, D, P.k, "cbd")
    e1 = CO.synth_small_np(P, O.TAG_E1, D, P.k, "uniform", P.error_bound_1)
    e2 = CO.synth_small_np(P, O.TAG_E2, D, P.n, "uniform", P.error_bound_2)
    return A, B, sk, m, r, e1, e2


def cpu_step(co, A, B, sk, m, r, e1, e2):
    c1, c2 = co.encrypt(A, B, m, r, e1, e2)
    return co.decrypt(sk, c1, c2)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import c_oracle as CO
    import pvw_oracle as O
    n = args.parties
    P = O.Params(n, K_DIM, ELL, p128_moduli())
    co = CO.COracle(P)
    D = args.cpu_dealers
    A, B, sk, m, r, e1, e2 = cpu_inputs(P, D)
    for _ in range(args.warmup):
        cpu_step(co, A, B, sk, m[:1], r[:1], e1[:1], e2[:1])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(co, A, B, sk, m, r, e1, e2)
    dt = time.perf_counter() - t0
    val = args.steps * D * n / dt
    sample = f"{D} dealers x {n} parties per step (encrypt + all-party decrypt), uniform synthetic B, {co.threads} OpenMP threads"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(n), "dealers_per_step": D, "note": "CPU port of the reference path (oracle/pvw_oracle.c); "
                       "the Rust crate and its fhe-math dependency cannot be built in this image"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": co.threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------------------------
def synth_device(torch, dev, gen, shape, kind, bound=None):
    if kind == "cbd":        # CBD(0.5): {-1, 0, 1}
        b = torch.randint(0, 4, shape, device=dev, generator=gen, dtype=torch.int64)
        return (b & 1) - ((b >> 1) & 1)
    if kind == "uniform":
        return torch.randint(-bound, bound + 1, shape, device=dev, generator=gen, dtype=torch.int64)
    if kind == "u63":
        return torch.randint(0, 2 ** 62, shape, device=dev, generator=gen, dtype=torch.int64)
    raise ValueError(kind)


def run_b200(args):
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
    n, k, l, L = args.parties, K_DIM, ELL, N_LIMBS
    plan = pvw.sharding.ShardPlan(n, world, rank)
    nrows, row0 = plan.nrows, plan.row0
    Dg = args.dealers                 # dealers per GPU per step (this rank's c1 slice)
    D = Dg * world                    # dealers per step
    c1_lo, c1_hi = plan.dealer_slice(D)
    moduli = p128_moduli()

    eng = pvw.Engine(n, k, l, moduli, row0=row0, nrows=nrows, device=local)
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    for kv in filter(None, os.environ.get("PVW_OPTS", "").split(",")):     # tuning knobs, e.g. PVW_OPTS=gemm_tile=0,refill_lag=1
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
    gen.manual_seed((SEED >> 8) & 0x7FFFFFFF)       # identical key material on every rank; each keeps its slice
    sk_all = synth_device(torch, dev, gen, (n, k, l), "cbd")
    sk = sk_all[row0:row0 + nrows].contiguous()
    del sk_all
    gen.manual_seed(1000 + rank)
    for p0 in range(0, nrows, 512):
        cnt = min(512, nrows - p0)
        ke = synth_device(torch, dev, gen, (cnt, k, l), "uniform", 100)
        eng.keygen_batch(row0 + p0, sk[p0:p0 + cnt].contiguous(), ke)
    eng.synchronize()
    eng.ct_reserve(D)

    # ---- the step's synthetic inputs: same dealers on every rank (r, e1), local columns of m / e2 ---------------
    gen.manual_seed(77)
    r = synth_device(torch, dev, gen, (D, k, l), "cbd")
    e1 = synth_device(torch, dev, gen, (D, k, l), "uniform", 100)
    gen.manual_seed(78 + rank)
    m = synth_device(torch, dev, gen, (D, nrows), "u63")
    e2 = synth_device(torch, dev, gen, (D, nrows, l), "uniform", 200)
    out = torch.empty((nrows, D), dtype=torch.int64, device=dev)
    parties = np.arange(row0, row0 + nrows, dtype=np.uint32)
    c1_view = eng.c1_store_tensor(0, D) if world > 1 else None

    def encrypt_and_gather(m_, r_, e1_, e2_):
        """c1 slice + c2 in one call, then the in-place all-gather of the c1 slices ordered on the library's stream.  (Issuing
        c1 first, its all-gather on a side stream and the c2 product meanwhile -- PVW_ENC_C1_ONLY / C2_ONLY -- was measured
        slower: 692.7 M against 738.6 M shares/s on 8 GPUs; the NCCL kernel does not co-reside with the persistent product.)"""
        eng.encrypt_batch(0, m_, r_, e1_, e2_, c1_range=(c1_lo, c1_hi))
        if world > 1:
            with torch.cuda.stream(ext):
                pvw.sharding.all_gather_c1(c1_view, plan)

    def step_device():
        encrypt_and_gather(m, r, e1, e2)
        eng.decrypt_batch(parties, sk, D=D, out=out)

    # host-buffer path (e2e): pinned host inputs, H2D inside the library calls, plaintexts read back to the host
    pin = lambda t: t.cpu().pin_memory()
    h_m, h_r, h_e1, h_e2, h_sk = pin(m), pin(r), pin(e1), pin(e2), pin(sk)
    n_m, n_r, n_e1, n_e2, n_sk = (t.numpy() for t in (h_m, h_r, h_e1, h_e2, h_sk))
    n_m = n_m.view(np.uint64)

    h_out = torch.empty((nrows, D), dtype=torch.int64).pin_memory()
    n_out = h_out.numpy().view(np.uint64)

    def step_host():
        encrypt_and_gather(n_m, n_r, n_e1, n_e2)
        return eng.decrypt_batch(parties, n_sk, D=D, out=n_out)

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
    step_device()
    eng.synchronize()
    bad_dev = (out.t() != m)
    res_host = step_host()
    bad_host = torch.from_numpy(res_host.view(np.int64).T != n_m.view(np.int64))
    for name, bad in (("device-resident", bad_dev), ("host-buffer", bad_host)):
        if bool(bad.any().item()):
            idx = bad.nonzero()
            raise SystemExit(f"bench: rank {rank}: {idx.shape[0]} decrypted shares differ from the messages on the {name} path "
                             f"(first (dealer, party) = {idx[0].tolist()}, dealers hit: {idx[:, 0].unique().numel()}, parties hit: "
                             f"{idx[:, 1].unique().numel()}) -- refusing to report a number")

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3) - 1):
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
    h2d = sum(a.nbytes for a in (n_m, n_r, n_e2, n_sk)) + n_e1[c1_lo:c1_hi].nbytes + parties.nbytes
    d2h = nrows * D * 8

    # ---- the reference's own call granularity: ONE encrypt / ONE decrypt_party_shares-per-party pass (D = 1).  The MAC kernel
    # is then a matrix-vector product that must stream B (or the secret keys) from HBM once: the HBM-bound case of the path.
    single = None
    if rank == 0:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
        single = {}
        for name, fn in (("encrypt", lambda: eng.encrypt_batch(0, m[:1], r[:1], e1[:1], e2[:1], c1_range=(0, 1))),
                         ("decrypt_all_parties", lambda: eng.decrypt_batch(parties, sk, D=1, out=out[:, :1].contiguous()))):
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
    mac_ms, mac_n, mac_bytes = prof["mac_gemm"]
    achieved = mac_bytes / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0
    kernel_ms = {kname: round(v[0] / args.steps, 4) for kname, v in prof.items()}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "mac_gemm_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    int_peak = None
    try:
        for ln in open(os.path.join(ROOT, "profiles", "r01_int_peaks.json")):
            if "mac_karatsuba_per_s" in ln:
                int_peak = float(json.loads(ln)["mac_karatsuba_per_s"])
    except Exception:
        pass
    mac_rate = (mac_bytes / ((k + 1.0) * 8)) * k / (mac_ms * 1e-3) if mac_ms > 0 else 0.0
    # NTT kernel against the measured Shoup modular-multiply rate: forward butterflies + gadget multiplies per step
    shoup_peak = None
    try:
        for ln in open(os.path.join(ROOT, "profiles", "r01_int_peaks.json")):
            if "mulmod_shoup_per_s" in ln:
                shoup_peak = float(json.loads(ln)["mulmod_shoup_per_s"])
    except Exception:
        pass
    ntt_ms, ntt_n, _ = prof["ntt_small"]
    polys = args.steps * (D * k + (c1_hi - c1_lo) * k + D * nrows + nrows * k)          # r, e1 slice, e2 (+m), sk
    log2l = l.bit_length() - 1
    ntt_mulmods = polys * L * ((l // 2) * log2l) + args.steps * D * nrows * L * l        # butterflies + m * g_hat
    ntt_rate = ntt_mulmods / (ntt_ms * 1e-3) if ntt_ms > 0 else 0.0
    # The batched product runs on the INT8 tensor cores (csrc/imma.cu): 64 u8 x u8 multiply-accumulates per 62-bit one, so the
    # kernel's roof is the tensor pipe.  Peak: kind::i8 issues K = 32 per instruction where bf16 issues K = 16 at the same
    # cadence, i.e. twice the dense bf16 rate.  MEASURED_PEAKS.json holds the measured cuBLAS bf16 figures; the product runs in
    # bursts of <= 2 ms between CUDA-core kernels (tensor duty 43 % of the step), so the burst figure is the denominator and the
    # fractions against the sustained figure and the nominal 4.5 POP/s are reported next to it.
    imma_on = os.environ.get("PVW_OPTS", "").replace(" ", "").find("imma=0") < 0
    bf16 = float(peaks.get("bf16_tflops", 2250.0))
    int8_peak = 2.0 * bf16
    int8_peak_sustained = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 2250.0)))
    int8_ops = mac_rate * 64 * 2 / 1e12
    hbm_view = {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "what": "SURVEY 8d algorithmic bytes (one operand row read per (dealer, row) + one polynomial written) over the kernel "
                        "time: > 1 because dealers are batched -- every staged tile serves 32 dealers from shared memory"}
    if imma_on:
        roofline = {"bound": "tensor", "kernel": "mac_gemm (imma_gemm_kernel: tcgen05.mma kind::i8)", "achieved": int8_ops, "peak": int8_peak,
                    "unit": "TOP/s (u8 x u8 -> s32, dense)", "frac": int8_ops / int8_peak,
                    "peak_source": ("2 x measured cuBLAS bf16 burst (MEASURED_PEAKS.json bf16_tflops)" if peaks else "2 x nominal dense bf16 (fallback)"),
                    "frac_of_2x_sustained_bf16": int8_ops / int8_peak_sustained, "frac_of_nominal_4500_TOPs": int8_ops / 4500.0,
                    "ncu_tensor_pipe_pct_of_peak": "63 % (sm__ops_path_tensor_op_utcimma_src_int8, c2 launch, profiles/r01_imma_gemm_ncu_summary.json)",
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
                        "what": "forward butterflies + gadget multiplies of r, e1, e2, sk per step over the kernel's event time; peak = "
                                "register-resident mulmod_shoup loop (profiles/r01_int_peaks.json)"},
                "single_call": {"what": "D = 1 (one reference-style encrypt call / one all-party decrypt pass): HBM-bound matrix-vector "
                                        "form on the CUDA cores (mac.cu), L2 flushed between calls, rows = %d" % nrows,
                                **{kname: dict(v, frac=v["achieved_GBps"] / peak) for kname, v in (single or {}).items()}},
                "modmuladds_per_s": mac_rate,
                "integer_pipe": {"achieved": mac_rate, "peak": int_peak, "unit": "62-bit modular multiply-accumulates/s",
                                 "frac": (mac_rate / int_peak) if int_peak else None,
                                 "peak_source": "ceiling of the CUDA-core form (3 IMAD.WIDE + carries per MAC, csrc/tools/int_peaks.cu, "
                                                "profiles/r01_int_peaks.json): the tensor-core kernel is measured against it for scale"},
                "note": "batched encrypt / decrypt multiply on the INT8 tensor cores: operands are byte planes, the 64 byte products of a "
                        "62-bit multiply accumulate on overlapping windows of the TMEM accumulator into 15 diagonal sums, recombined and "
                        "reduced exactly in the epilogue (DESIGN.md 4); PVW_OPTS=imma=0 runs the CUDA-core kernel instead"})

    # ---- CPU baseline: oracle port on the host cores, bounded sample -------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import c_oracle as CO
        import pvw_oracle as O
        P = O.Params(n, k, l, moduli, psi=eng.psi)
        co = CO.COracle(P)
        Dc = min(args.cpu_baseline_dealers, Dg)
        A_h = eng.crs_download()
        B_h = eng.pk_download_rows(0, n)
        r_h, e1_h, e2_h, m_h, sk_h = n_r[:Dc], n_e1[:Dc], n_e2[:Dc], n_m[:Dc], n_sk
        t0 = time.perf_counter()
        c1_h, c2_h = co.encrypt(A_h, B_h, m_h, r_h, e1_h, e2_h)
        dec_h = co.decrypt(sk_h, c1_h, c2_h)
        dt = time.perf_counter() - t0
        # the CPU port and the GPU agree on this sample (same keys, same randomness): plaintexts and ciphertext of dealer 0
        g1, g2 = eng.ct_download(0)
        same = bool((dec_h == res_host[:, :Dc]).all() and (g1 == c1_h[0]).all() and (g2 == c2_h[0]).all())
        cpu = {"value": Dc * n / dt, "unit": UNIT, "cores": co.threads, "kind": "port",
               "sample": f"{Dc} dealers x {n} parties (encrypt + all-party decrypt) of the same workload, {dt:.1f} s; "
                         f"bit-identical to the GPU result: {same}"}
        if not same:
            raise SystemExit("bench: CPU port and GPU disagree on the sample")

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(n), "dealers_per_step": D, "dealers_per_gpu_per_step": Dg, "shares_per_step": shares_per_step,
                       "parallelism": f"B rows / parties sharded x{world}; c1 dealer slices all-gathered (NCCL)" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (B shard %.0f MB, ciphertext store %.0f MB per step vs 126 MB L2)" % (
                           nrows * k * L * l * 8 / 1e6, D * (nrows + k) * L * l * 8 / 1e6),
                       "keys": "genuine (device keygen); every decrypted share checked == message before timing"},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "roofline": roofline, "cpu_baseline": cpu,
            "hbm_roof_shares_per_s": peak * 1e9 / bytes_per_share(n) * world}
    emit(line)
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
    ap.add_argument("--dealers", type=int, default=256, help="dealers per GPU per step")
    ap.add_argument("--parties", type=int, default=N_PARTIES)
    ap.add_argument("--cpu-dealers", type=int, default=16, help="dealers per step of the --impl reference arm")
    ap.add_argument("--cpu-baseline-dealers", type=int, default=64, help="dealers in the cpu_baseline sample (about 11 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
