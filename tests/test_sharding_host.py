"""Host-side logic of the row-block sharded path (imageclust_b200/sharding.py) on the CPU:
row-range arithmetic, and the handle all-gather over a world_size-2 ``gloo`` group."""
import os
import socket
import sys

import numpy as np
import pytest

from imageclust_b200 import _lib, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,world", [(0, 2), (1, 8), (7, 2), (8, 8), (1000, 3), (100_000, 8), (250_000, 8), (20_001, 4)])
def test_row_ranges_partition_the_slots(n, world):
    c = sharding.rows_per_rank(n, world)
    assert c % 4 == 0 and c * world >= n and (n == 0 or (c - 4) * world < n)
    covered = []
    for r in range(world):
        lo, hi = sharding.row_range(n, r, world)
        assert 0 <= lo <= hi <= n and hi - lo <= c
        covered.extend(range(lo, hi))
        for s in (lo, hi - 1):
            if lo < hi:
                assert sharding.owner_of(s, n, world) == r
    assert covered == list(range(n))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from imageclust_b200 import _lib as L, sharding as S

    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeEngine:  # records what the sharded wrapper asks the C ABI to do
        def __init__(self):
            self.calls = []

        def shard_init(self, r, w):
            self.calls.append(("init", r, w))

        def load(self, x):
            self.calls.append(("load", x.shape))

        def shard_export(self):
            return bytes([rank + 1]) * L.SHARD_HANDLE_BYTES

        def shard_connect(self, blobs):
            self.calls.append(("connect", [b[0] for b in blobs], [len(b) for b in blobs]))

        def run_resident(self, mn, mx):
            self.calls.append(("run", mn, mx))
            return "result"

    eng = FakeEngine()
    sh = S.ShardedEngine(eng, rank, world)
    x = np.zeros((10, 4), np.float32)
    sh.load(x)
    sh.load(x)  # same shape: no second handle exchange
    out = sh.run_resident(2, 5)
    sh.load(np.zeros((12, 4), np.float32))  # new shape: reconnect
    q.put((rank, eng.calls, out))
    dist.destroy_process_group()


def test_handle_exchange_over_gloo_world_size_2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, calls, out in got:
        assert out == "result"
        assert calls[0] == ("init", rank, 2)
        connects = [c for c in calls if c[0] == "connect"]
        assert len(connects) == 2  # once per distinct shape
        for c in connects:
            assert c[1] == [1, 2] and c[2] == [_lib.SHARD_HANDLE_BYTES] * 2  # every rank's blob, in rank order
        assert [c[0] for c in calls] == ["init", "load", "connect", "load", "run", "load", "connect"]
