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


def test_reference_arm_contract_under_torchrun(built):
    """bench.py --impl reference (the reference's own CPU code on the host cores, no GPU involved) launched the way the
    driver launches every arm for N > 1: rank 0 alone prints ONE JSON line, the other rank exits 0 without work; the line
    names the same configuration the repository's arm would name for the same flags."""
    import json
    import subprocess
    port = 29600 + (os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--scans", "8"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scans_per_sec" and d["n_gpus"] == 2 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, str(ROOT))
    import bench
    assert d["config"] == bench.config_dict(8, 2)
