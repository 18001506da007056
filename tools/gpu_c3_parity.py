"""Parity at BASELINE configs[2] scale: the stream of 4,096 C2 scans (seeds 1000..5095) sharded by frame over the ranks
(torchrun, one rank per GPU, 512 scans per GPU at 8 GPUs), every scan compared with the CPU oracle on the host cores:
ring/sector keys exactly, labels for the default solver and for the reference-order mode.  Rank 0 writes the table.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/gpu_c3_parity.py [total_scans] [out.md]
"""
import importlib, os, sys, time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import oracle_lib
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
total_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
out_path = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / "parity_c3.md"
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")  # control plane only: the table is gathered on the host, no data-path collective
own = rpw.sharding.shard_range(total_scans, rank, world)
lo, hi = own.start, own.stop
threads = max(1, (os.cpu_count() or 8) // world)
cfg = rpw.PatchworkConfig(filtering_radius=80.0)
ocfg = oracle_lib.to_cfg(cfg)
oracle = oracle_lib.Oracle()
t0 = time.perf_counter()
rows = []
CH = 128  # scans per call: bounds host memory (the oracle's outputs) and keeps the batch kernels busy
with ThreadPoolExecutor(threads) as ex:
    h = rpw.Handle(cfg.to_c(), local, CH * 120000 + 4096, CH)
    for c0 in range(lo, hi, CH):
        seeds = range(1000 + c0, 1000 + min(hi, c0 + CH))
        scans = list(ex.map(lambda s: rpw.synth.spinning_scan(int(s)), seeds))
        want = list(ex.map(lambda a: oracle.run(ocfg, a), scans))
        n = [len(a) for a in scans]
        h.set_plane_solver(rpw.capi.SOLVER_HYBRID)
        fast = h.segment_batch(scans)
        keys = h.debug_keys(sum(n))
        h.set_plane_solver(rpw.capi.SOLVER_REFERENCE)
        exact = h.segment_batch(scans)
        o = 0
        for s, a, w, lf, le in zip(seeds, scans, want, fast, exact):
            rows.append((s, len(a), int((keys[o:o + len(a)] != w["keys"]).sum()), int((lf != w["labels"]).sum()), int((le != w["labels"]).sum())))
            o += len(a)
    h.close()
dt = time.perf_counter() - t0
arr = np.array(rows, np.int64).reshape(-1, 5)
if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, arr)
    arr = np.concatenate(gathered)
    dist.barrier()
if rank == 0:
    pts = int(arr[:, 1].sum())
    worst_fast = float((1 - arr[:, 3] / arr[:, 1]).min()); worst_exact = float((1 - arr[:, 4] / arr[:, 1]).min())
    below = arr[arr[:, 3] > 0.001 * arr[:, 1]]
    lines = [f"# Parity at BASELINE configs[2] scale: {len(arr)} C2 scans (seeds 1000..{999 + len(arr)}), {pts} points, sharded by frame over {world} GPU(s)", "",
             f"Every scan against the CPU oracle (pinned bit for bit to the reference's strict build), oracle on {threads} host threads per rank, {dt:.0f} s wall on rank 0.", "",
             "| | default solver (hybrid) | reference-order mode (RPW_SOLVER_REFERENCE) |", "|---|---|---|",
             f"| ring/sector key mismatches | {int(arr[:, 2].sum())} of {pts} | (same keys) |",
             f"| labels differing | {int(arr[:, 3].sum())} of {pts} ({1 - arr[:, 3].sum() / pts:.8f} agreement) | {int(arr[:, 4].sum())} of {pts} |",
             f"| scans with any difference | {int((arr[:, 3] > 0).sum())} | {int((arr[:, 4] > 0).sum())} |",
             f"| worst scan agreement | {worst_fast:.6f} | {worst_exact:.6f} |",
             f"| scans below the 99.9 % bar | {len(below)} | {int((arr[:, 4] > 0.001 * arr[:, 1]).sum())} |", ""]
    if len(below):
        lines += ["Scans below the bar with the default solver (seed, points, labels differing): " + ", ".join(f"{int(r[0])} ({int(r[1])}, {int(r[3])})" for r in below), ""]
    worst = arr[np.argsort(-arr[:, 3])[:8]]
    lines += ["Largest differences, default solver (seed: labels differing): " + ", ".join(f"{int(r[0])}: {int(r[3])}" for r in worst if r[3] > 0)]
    out_path.parent.mkdir(parents=True, exist_ok=True)
    out_path.write_text("\n".join(lines) + "\n")
    print("\n".join(lines))
if world > 1:
    dist.destroy_process_group()
