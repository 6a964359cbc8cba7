"""CPU: host-side logic of the package that needs no device -- the CSPRNG-backed samplers (the reference samples from
thread_rng(), a CryptoRng: src/keys/secret_key.rs:45-63, src/crypto/encryption.rs:138,164,180), ciphertext-slot pinning in
decrypt_party_shares (src/crypto/decryption.rs:281-325), input validation, and the handle exchange of the multi-GPU c1
exchange over gloo with a fake engine."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import pvw_rs_b200 as pvw                      # noqa: E402  (imports without a device; only Engine() needs one)
from pvw_rs_b200 import api                    # noqa: E402


# ---- sampling ---------------------------------------------------------------------------------------------------
def test_default_generator_is_the_os_csprng():
    assert isinstance(api._rng(None), api._OsRng)
    g = np.random.default_rng(1)
    assert api._rng(g) is g                      # explicit generators are for reproducible tests only


def test_os_rng_ranges_and_moments():
    g = api._OsRng()
    x = g.integers(-200, 201, size=200000, dtype=np.int64)
    assert x.min() == -200 and x.max() == 200
    assert abs(x.mean()) < 2.0 and abs(x.var() - (401 ** 2 - 1) / 12) < 400
    y = g.integers(-1172385, 1172386, size=1000, dtype=np.int64)          # examples/pvw_valid_dec.rs:52
    assert y.min() >= -1172385 and y.max() <= 1172385 and len(np.unique(y)) > 990
    q = 0x3FFFFFFFFFFFFDC1
    z = g.integers(0, q, size=(3, 4), dtype=np.uint64)
    assert z.dtype == np.uint64 and z.shape == (3, 4) and int(z.max()) < q and int(z.max()) > (1 << 40)
    assert len({int(g.integers(0, 1 << 62)) for _ in range(8)}) == 8       # not a constant stream


def test_cbd_shapes_match_the_reference():
    """src/sampling/uniform.rs:27-70 and tests/sampling.rs:197-274: variance 0.5 -> {-1, 0, 1}, mean ~ 0, variance ~ 0.5"""
    x = pvw.sample_vec_cbd(10000, 0.5)
    assert set(np.unique(x)) <= {-1, 0, 1}
    assert abs(x.mean()) < 0.05 and abs(x.var() - 0.5) < 0.05
    y = pvw.sample_vec_cbd(20000, 10.0)
    assert np.abs(y).max() <= 20 and abs(y.var() - 10.0) < 0.6
    with pytest.raises(pvw.PvwError):
        pvw.sample_vec_cbd(4, 0.0)
    with pytest.raises(pvw.PvwError):
        pvw.sample_vec_cbd(4, 2.5)
    u = pvw.sample_uniform_coefficients(100, 5000)
    assert u.min() >= -100 and u.max() <= 100 and u.dtype == np.int64


def test_scalars_outside_u64_are_rejected():
    assert api._as_scalars([0, 1, (1 << 64) - 1], "x").dtype == np.uint64
    for bad in ([-1], [1 << 64], ["a"]):
        with pytest.raises(pvw.PvwError) as ei:
            api._as_scalars(bad, "scalars")
        assert ei.value.variant == "InvalidParameters"


# ---- ciphertext slots -------------------------------------------------------------------------------------------
class _FakeEngine:
    """records uploads / downloads; slot contents are (c1, c2) pairs"""

    def __init__(self):
        self.capacity, self.slots, self.decrypted = 0, {}, None

    def ct_reserve(self, cap):
        self.capacity = cap

    def ct_upload(self, slot, c1=None, c2=None):
        self.slots[slot] = (c1, c2)

    def ct_download(self, slot, want_c1=True, want_c2=True):
        c1, c2 = self.slots[slot]
        return (c1 if want_c1 else None), (c2 if want_c2 else None)

    def decrypt_batch(self, party_idx, sk, dealer_slots=None, D=None, out=None):
        self.decrypted = list(dealer_slots)
        return np.array([[int(self.slots[s][0][0]) for s in dealer_slots]], dtype=np.uint64)


class _FakeParams:
    n, k, l, L = 4, 1, 1, 1


class _FakePk:
    def __init__(self, cap):
        import threading
        self.params, self.engine = _FakeParams(), _FakeEngine()
        self._slots = api._SlotPool(self.engine, cap)
        self._lock = threading.RLock()


class _FakeSk:
    secret_coeffs = np.zeros((1, 1), dtype=np.int64)


def test_decrypt_party_shares_never_evicts_a_ciphertext_of_the_same_call():
    """capacity == n with mixed resident and spilled inputs: Y0 resident, the pool then fills up with other
    ciphertexts, Y1..Y3 are host-side only -- every dealer must be decrypted from its own data"""
    pk = _FakePk(4)
    ys = [api.PvwCiphertext(pk, c1=np.array([100 + d], dtype=np.uint64), c2=np.zeros(4, dtype=np.uint64)) for d in range(4)]
    ys[0]._resident_slot()                                                   # an earlier decrypt_party_value made Y0 resident
    others = [api.PvwCiphertext(pk, c1=np.array([900 + i], dtype=np.uint64), c2=np.zeros(4, dtype=np.uint64)) for i in range(3)]
    for o in others:
        o._resident_slot()                                                   # the pool is now full: Y0 + three strangers
    got = api.decrypt_party_shares(ys, _FakeSk(), 2)
    assert got == [100, 101, 102, 103]
    assert len(set(pk.engine.decrypted)) == 4                                # four distinct slots
    assert all(o._slot is None for o in others) and [int(o.c1[0]) for o in others] == [900, 901, 902]   # strangers spilled intact


def test_slot_pool_refuses_when_everything_is_pinned():
    pk = _FakePk(1)
    a = api.PvwCiphertext(pk, c1=np.array([1], dtype=np.uint64), c2=np.zeros(4, dtype=np.uint64))
    b = api.PvwCiphertext(pk, c1=np.array([2], dtype=np.uint64), c2=np.zeros(4, dtype=np.uint64))
    s = a._resident_slot()
    with pytest.raises(pvw.PvwError):
        b._resident_slot(pinned={s})


# ---- narrow inputs ----------------------------------------------------------------------------------------------
def test_small_input_element_types_pick_the_flags():
    from pvw_rs_b200 import _ffi, engine
    a = engine._SmallArg(np.zeros((2, 3, 8), dtype=np.int8), (2, 3, 8), "r", "secret")
    assert a.flag == _ffi.PVW_IN_SECRET_I8 and a.bytes == 1
    b = engine._SmallArg(np.zeros((2, 3, 8), dtype=np.int32), (2, 3, 8), "e", "error")
    assert b.flag == _ffi.PVW_IN_ERROR_I32
    c = engine._SmallArg(np.zeros((2, 3, 8), dtype=np.int16), (2, 3, 8), "e", "error")
    assert c.flag == _ffi.PVW_IN_ERROR_I16
    d = engine._SmallArg(np.zeros((2, 3, 8), dtype=np.int32), (2, 3, 8), "r", "secret")     # not a secret type: widened on the host
    assert d.flag == 0 and d.bytes == 8
    e = engine._SmallArg([[[1] * 8] * 3] * 2, (2, 3, 8), "r", "secret")
    assert e.flag == 0 and e.keep.dtype == np.int64
    with pytest.raises(pvw.PvwError):
        engine._SmallArg(np.zeros((2, 3, 7), dtype=np.int8), (2, 3, 8), "r", "secret")
    with pytest.raises(pvw.PvwError):
        engine._same_error_type(b, c)


# ---- multi-GPU exchange: host protocol over gloo ------------------------------------------------------------------
class _FakeShardEngine:
    def __init__(self, rank):
        self.rank, self.calls = rank, []

    def shard_export(self, world):
        return bytes([self.rank, world]) + bytes(190)

    def shard_connect(self, world, rank, handles):
        self.calls.append(("connect", world, rank, [h[0] for h in handles], [len(h) for h in handles]))

    def shard_push_c1(self, slot0, count):
        self.calls.append(("push", slot0, count))

    def shard_wait_c1(self):
        self.calls.append(("wait",))

    def shard_release_c1(self):
        self.calls.append(("release",))

    def shard_disconnect(self):
        self.calls.append(("disconnect",))


def _exchange_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pvw_rs_b200 as pkg
        plan = pkg.sharding.ShardPlan(13, world, rank)
        eng = _FakeShardEngine(rank)
        x = pkg.sharding.CopyEngineExchange(eng, plan)
        x.push(4, 6)
        x.wait()
        x.release()
        x.close()
        x.close()
        with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as f:
            f.write(repr(eng.calls))
    finally:
        dist.destroy_process_group()


def test_copy_engine_exchange_protocol_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29950 + os.getpid() % 40
    mp.spawn(_exchange_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        calls = eval(open(tmp_path / f"r{rank}.txt").read())
        assert calls[0] == ("connect", 2, rank, [0, 1], [192, 192])           # every rank got both handles, in rank order
        assert calls[1] == ("push", 4 + 3 * rank, 3)                          # its dealer slice of the 6 dealers stored from slot 4
        assert calls[2:] == [("wait",), ("release",), ("disconnect",)]        # close() is idempotent


def test_exchange_is_a_no_op_on_one_gpu():
    eng = _FakeShardEngine(0)
    x = pvw.sharding.CopyEngineExchange(eng, pvw.sharding.ShardPlan(7, 1, 0))
    x.push(0, 4)
    x.wait()
    x.release()
    x.close()
    assert eng.calls == []
