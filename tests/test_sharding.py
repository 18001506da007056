"""Frame sharding across ranks: host logic only, exercised with gloo, world size 2."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_shard_ranges_partition_the_stream(rpw):
    sh = rpw.sharding
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 3, 8):
            blocks = [list(sh.shard_range(n, r, w)) for r in range(w)]
            assert sorted(sum(blocks, [])) == list(range(n))
            assert max(map(len, blocks)) - min(map(len, blocks)) <= 1
            rr = [list(sh.shard_round_robin(n, r, w)) for r in range(w)]
            assert sorted(sum(rr, [])) == list(range(n))


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    import importlib
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
    n_scans = 7
    owners = [list(rpw.sharding.shard_range(n_scans, r, world)) for r in range(world)]
    # stand-in for per-scan label buffers: every rank "segments" only its own frames
    local = [np.full(10 + f, f, np.uint8) for f in owners[rank]]
    out = rpw.sharding.gather_labels(local, owners, dist, dst=0)
    if rank == 0:
        ok = len(out) == n_scans and all(len(out[f]) == 10 + f and (out[f] == f).all() for f in range(n_scans))
        q.put(ok)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_label_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
