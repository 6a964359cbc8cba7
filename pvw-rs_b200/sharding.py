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
