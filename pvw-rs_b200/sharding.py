"""Row sharding of the public-key matrix B across the GPUs of one box.

Rank g owns rows [g*n/G, (g+1)*n/G) of B -- i.e. those parties' c2 rows and decryptions.  Within a step of D dealers
every rank computes c1 for its contiguous slice of the dealers and the slices are all-gathered (the only collective on
the path).  This module is pure host logic (no CUDA): the same plan drives NCCL on the GPUs and gloo in the CPU tests."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class ShardPlan:
    n: int          # parties = rows of B
    world: int
    rank: int

    def __post_init__(self):
        if self.world <= 0 or not 0 <= self.rank < self.world:
            raise ValueError("bad rank / world size")
        if self.n < self.world:
            raise ValueError(f"cannot shard n={self.n} parties over {self.world} ranks")

    def rows_of(self, rank: int) -> Tuple[int, int]:
        """(row0, nrows) of `rank`: the first n % world ranks hold one extra row"""
        base, extra = divmod(self.n, self.world)
        row0 = rank * base + min(rank, extra)
        return row0, base + (1 if rank < extra else 0)

    @property
    def row0(self) -> int:
        return self.rows_of(self.rank)[0]

    @property
    def nrows(self) -> int:
        return self.rows_of(self.rank)[1]

    def owner_of(self, party: int) -> int:
        if not 0 <= party < self.n:
            raise IndexError(party)
        base, extra = divmod(self.n, self.world)
        cut = extra * (base + 1)
        return party // (base + 1) if party < cut else extra + (party - cut) // base

    def dealer_slice(self, D: int, rank: int = None) -> Tuple[int, int]:
        """dealers [lo, hi) of a step whose c1 this rank computes; requires D % world == 0 for an in-place all-gather"""
        rank = self.rank if rank is None else rank
        if D % self.world:
            raise ValueError(f"{D} dealers per step do not divide over {self.world} ranks")
        per = D // self.world
        return rank * per, (rank + 1) * per


def all_gather_c1(c1_store, plan: ShardPlan, group=None):
    """In-place all-gather of the c1 dealer slices.  `c1_store` is a [D][words] tensor over the ciphertext store
    (Engine.c1_store_tensor on the GPU; any tensor with the same shape in the CPU tests)."""
    import torch.distributed as dist
    if plan.world == 1:
        return
    lo, hi = plan.dealer_slice(c1_store.shape[0])
    dist.all_gather_into_tensor(c1_store, c1_store[lo:hi], group=group)


class CopyEngineExchange:
    """The c1 exchange of a step without an all-gather kernel (include/pvw_b200.h, pvw_shard_*): every rank pushes its dealer
    slice into the peers' ciphertext stores with copy-engine peer copies over NVLink, ordered by stream counters -- no SM, no
    host synchronisation, so it runs under the c2 product.  This class is the host protocol only: it all-gathers the IPC
    handles once (any torch.distributed backend; bytes travel as a uint8 tensor) and forwards the per-step calls.
    `engine` needs shard_export / shard_connect / shard_push_c1 / shard_wait_c1 / shard_release_c1 / shard_disconnect."""

    HANDLE_BYTES = 192

    def __init__(self, engine, plan: ShardPlan, group=None, device=None):
        self.engine, self.plan, self.connected = engine, plan, False
        if plan.world == 1:
            return
        import torch
        import torch.distributed as dist
        mine = torch.frombuffer(bytearray(engine.shard_export(plan.world)), dtype=torch.uint8)
        if len(mine) != self.HANDLE_BYTES:
            raise ValueError("unexpected handle size")
        if device is not None:
            mine = mine.to(device)
        table = torch.empty(plan.world * self.HANDLE_BYTES, dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(table, mine, group=group)
        table = table.cpu().numpy().reshape(plan.world, self.HANDLE_BYTES)
        engine.shard_connect(plan.world, plan.rank, [row.tobytes() for row in table])
        self.connected = True

    def note_pushed(self):
        """the push was queued inside pvw_encrypt_batch (PVW_ENC_PUSH_C1): nothing to do here, kept for symmetry / call counting"""

    def push(self, slot0: int, D: int):
        """after the c1 product of this rank's slice of the D dealers stored from slot0 has been queued"""
        if self.plan.world == 1:
            return
        lo, hi = self.plan.dealer_slice(D)
        self.engine.shard_push_c1(slot0 + lo, hi - lo)

    def wait(self):
        if self.plan.world > 1:
            self.engine.shard_wait_c1()

    def release(self):
        if self.plan.world > 1:
            self.engine.shard_release_c1()

    def close(self):
        if self.connected:
            self.engine.shard_disconnect()
            self.connected = False
