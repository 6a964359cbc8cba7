"""CPU, world size 2 over gloo: the host-side sharding plan + the c1 all-gather reproduce the single-process result.
The oracle stands in for the kernels (each rank: c2 for its rows, c1 for its dealer slice), exactly the split
bench.py runs on the GPUs over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("pvw_sharding", os.path.join(ROOT, "pvw-rs_b200", "sharding.py"))
        sh = importlib.util.module_from_spec(spec)
        sys.modules["pvw_sharding"] = sh
        spec.loader.exec_module(sh)
        from _cases import System, params
        P = params("RAG")                       # n = 13 parties: uneven shards (7 + 6)
        D = 6
        S = System(P, D, "example")
        plan = sh.ShardPlan(P.n, world, rank)
        lo, hi = plan.row0, plan.row0 + plan.nrows
        dlo, dhi = plan.dealer_slice(D)
        # local work: c2 for my rows x all dealers, c1 only for my dealer slice
        _, c2 = S.co.encrypt(S.A, S.B[lo:hi], S.m[:, lo:hi], S.r, S.e1, S.e2[:, lo:hi], want_c1=False)
        c1_mine, _ = S.co.encrypt(S.A, S.B[:1], S.m[dlo:dhi, :1], S.r[dlo:dhi], S.e1[dlo:dhi], S.e2[dlo:dhi, :1], want_c2=False)
        words = P.k * P.L * P.l
        store = torch.zeros((D, words), dtype=torch.int64)
        store[dlo:dhi] = torch.from_numpy(c1_mine.reshape(dhi - dlo, words).view(np.int64))
        sh.all_gather_c1(store, plan)
        c1 = store.numpy().view(np.uint64).reshape(D, P.k, P.L, P.l)
        dec = S.co.decrypt(S.sk[lo:hi], c1, c2)                 # [nrows][D]
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, dec))
        if rank == 0:
            full = np.zeros((P.n, D), dtype=np.uint64)
            for a, b, d in gathered:
                full[a:b] = d
            c1_ref, c2_ref = S.encrypt()
            ok = bool((c1 == c1_ref).all() and (full == S.co.decrypt(S.sk, c1_ref, c2_ref)).all() and (full == S.m.T).all())
            with open(out_path, "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


def test_two_rank_row_sharding_over_gloo(tmp_path):
    out = str(tmp_path / "result.txt")
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shard_plan_arithmetic():
    import importlib.util
    spec = importlib.util.spec_from_file_location("pvw_sharding", os.path.join(ROOT, "pvw-rs_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec)
    sys.modules["pvw_sharding"] = sh
    spec.loader.exec_module(sh)
    for n, world in [(13, 2), (4096, 8), (7, 7), (10, 4), (8192, 8)]:
        covered = []
        for r in range(world):
            p = sh.ShardPlan(n, world, r)
            covered += list(range(p.row0, p.row0 + p.nrows))
            for q in (p.row0, p.row0 + p.nrows - 1):
                assert p.owner_of(q) == r
        assert covered == list(range(n))
    p = sh.ShardPlan(4096, 8, 3)
    assert (p.row0, p.nrows) == (1536, 512) and p.dealer_slice(256) == (96, 128)
    with pytest.raises(ValueError):
        sh.ShardPlan(4, 8, 0)
    with pytest.raises(ValueError):
        p.dealer_slice(100)
