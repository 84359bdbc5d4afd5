"""Row-block sharding of ONE clustering over the GPUs of a box: host-side plumbing.

One process per GPU (``torchrun``), one :class:`~imageclust_b200.clustering.Engine` per
process.  The CUDA library does the work (``ic_shard_*`` in ``include/imageclust_b200.h``):
every rank keeps the rows of its slot block with all their columns, the ranks' persistent
merge-loop kernels exchange one 128-byte record per merge through peer-mapped memory
(CUDA IPC over NVLink), and rows ``a`` / ``b`` of a merge are read from / written to their
owner directly.  The host only has to move the ranks' IPC handles around once per problem
shape -- that is all this module uses ``torch.distributed`` for.

The reference has no counterpart (its ``[][]float32`` matrix lives in one Go process,
``/root/reference/internal/clustering/clustering.go:61-73``); the sharding is BASELINE.json's
north_star ("row-block sharded across the 8 GPUs of one box").
"""
from __future__ import annotations

from . import _lib


def rows_per_rank(n: int, world: int) -> int:
    """C = ceil(n / world) rounded up to a multiple of 4: rank r owns the slots [r*C, (r+1)*C)."""
    return (-(-int(n) // int(world)) + 3) // 4 * 4 if n > 0 else 0


def row_range(n: int, rank: int, world: int):
    """[row_begin, row_end) of ``rank`` -- the same arithmetic as ``ic_shard_rows``."""
    c = rows_per_rank(n, world)
    lo = min(n, rank * c)
    return lo, min(n, lo + c)


def owner_of(slot: int, n: int, world: int) -> int:
    return slot // rows_per_rank(n, world)


def gather_blobs(blob: bytes, group=None, device=None):
    """All-gather one fixed-size byte blob per rank (rank order) with torch.distributed."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


class ShardedEngine:
    """An :class:`Engine` that is rank ``rank`` of ``world`` row-block shards.

    ``load`` / ``cluster`` / ``run_resident`` are collective: every rank calls them with the
    same matrix and constraints, every rank gets the complete result."""

    def __init__(self, engine, rank: int, world: int, group=None):
        self.eng = engine
        self.rank, self.world, self.group = int(rank), int(world), group
        engine.shard_init(rank, world)
        self._shape = None

    def _connect(self):
        blobs = gather_blobs(self.eng.shard_export(), self.group)
        assert len(blobs) == self.world and all(len(b) == _lib.SHARD_HANDLE_BYTES for b in blobs)
        self.eng.shard_connect(blobs)

    def load(self, x):
        self.eng.load(x)
        if self._shape != tuple(x.shape):  # same shape: the library keeps its allocations and peer mappings
            self._connect()
            self._shape = tuple(x.shape)

    def run_resident(self, min_size: int, max_size: int):
        return self.eng.run_resident(min_size, max_size)

    def cluster(self, x, min_size: int, max_size: int):
        """Upload + run.  The first call for a shape connects the ranks (load, exchange handles)."""
        if self._shape != tuple(x.shape):
            self.load(x)
        return self.eng.cluster(x, min_size, max_size)
