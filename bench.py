#!/usr/bin/env python
"""bench.py — scans/s of the ground-segmentation hot path on B200(s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1] streamed as configs[2]): KITTI-like synthetic 64-beam spinning
scans, 120,000 points each, R = 80 m, default zone model, seeds 1000+.  A *step* is one pass of
the whole path (bin -> offsets -> scatter -> fit/label) over a batch of `--scans` distinct scans;
with N GPUs every rank owns its own batch (frames shard across GPUs, no data-path collective,
weak scaling).  `value` = scans all ranks processed / max-over-ranks device time, with the batch
already resident in HBM (batch bytes > L2, so no L2 carry-over between steps).  `e2e` = the same
metric through the C-ABI with pinned HOST buffers, host->device and device->host copies inside the
timed region, beside the copy ceiling the ranks measure together (`e2e.ceiling_scans_per_sec`).
`roofline` is for the dominant kernel (fit), timed live with CUDA events on the launching stream;
`roofline_kernels` carries bin and scatter the same way.  `cpu_baseline` / `--impl reference` time
the reference's own CPU implementation (oracle/_ref/libref_fast.so, the reference's translation
units built with its library flags) on the box's host cores.  `shapes` repeats the measurement for
the other named shapes (C1 test-suite cloud, C4 merged solid-state frame, C5 dense urban scan).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import struct
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
PKG = "ros2-recursive-patchwork-implementation_b200"

POINTS_PER_SCAN = 120000
ALG_BYTES_PER_POINT = 17  # 16 B float4 read + 1 B label write
PASS_BYTES = {"bin": 18, "scatter": 34}  # pass-model bytes per point: K1 16 R + 2 W (key); K2 2 R + 16 R + 16 W
METRIC = "scans_per_sec"
UNIT = "scans/s (120k-point 64-beam scans)"
WORKLOAD = "C2 KITTI-like 64-beam 120k-point scans (BASELINE configs[1]) streamed as configs[2] batches, R=80 m, default zone model, seeds 1000+"


def config_dict(scans_per_step, world):
    """The same for both arms (the driver compares them): the workload, not how an arm samples it."""
    return {"workload": WORKLOAD, "scans_per_step_per_gpu": scans_per_step, "points_per_scan": POINTS_PER_SCAN,
            "parallelism": f"frames x{world}",
            "l2": f"batch is {scans_per_step * POINTS_PER_SCAN * 16 / 1e6:.0f} MB of float4 input per GPU > 126 MB L2 (no flush needed)"}


def gen_scans(rpw, seeds, threads, gen=None):
    gen = gen or (lambda s: rpw.synth.spinning_scan(int(s)))
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(gen, seeds))


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "samples": len(sm), "reasons": sorted(reasons)}


def load_reference():
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib
    ref = oracle_lib.try_reference("fast")
    if ref is not None:
        return ref, "reference", oracle_lib
    return oracle_lib.Oracle(), "port", oracle_lib  # the C restatement when oracle/_ref was not built


def cpu_reference_throughput(scans, cfg, seconds_budget, threads, what="C2 seeds"):
    """Reference CPU implementation on the host cores: frame-parallel, one RecursivePatchwork per call
    (the class is stateless, RP/include/recursive_patchwork.hpp:70).  Returns scans/s and a description."""
    ref, kind, oracle_lib = load_reference()
    ccfg = oracle_lib.to_cfg(cfg)
    t_one = ref.time_scan(ccfg, scans[0], 1)  # warm-up + estimate
    n_jobs = max(threads, min(4000, int(seconds_budget / max(t_one, 1e-4))))  # ~seconds_budget CPU-seconds of work in total
    jobs = [scans[i % len(scans)] for i in range(n_jobs)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda a: ref.time_scan(ccfg, a, 1), jobs))
    dt = time.perf_counter() - t0
    return n_jobs / dt, kind, f"{n_jobs} scans ({len(scans)} distinct, {what}) on {threads} threads, {dt:.1f} s wall, libref_{'fast' if kind == 'reference' else 'oracle'}"


def run_reference(args, emit):
    """The reference arm: the reference's own CPU implementation of the path, all host threads, frame-parallel, on the
    SAME configuration (workload, scans per step, seeds) as the repository's arm; rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rpw = importlib.import_module(PKG)
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    threads = os.cpu_count() or 1
    B = args.scans
    scans = gen_scans(rpw, range(1000, 1000 + B), threads)
    ref, kind, oracle_lib = load_reference()
    ccfg = oracle_lib.to_cfg(cfg)
    # a step = a bounded sample of the step's 512 scans (every k-th scan of the batch), sized so that warm-up + steps
    # stay within about two minutes at ~50 ms per scan and thread
    t_one = ref.time_scan(ccfg, scans[0], 1)
    budget_scans = max(threads * 4, int(100.0 * threads / max(t_one, 1e-4) / max(1, args.steps + args.warmup)))
    per_step = min(B, budget_scans)
    pick = [scans[int(i * B / per_step)] for i in range(per_step)]

    with ThreadPoolExecutor(max_workers=threads) as ex:
        def step():
            list(ex.map(lambda a: ref.time_scan(ccfg, a, 1), pick))

        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    sample = f"{per_step} of the step's {B} scans per step on {threads} host threads (libref_{'fast' if kind == 'reference' else 'oracle'})"
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps * (B / per_step), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(B, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mpoints_per_sec": val * POINTS_PER_SCAN / 1e6,
    })


def resident_rate(rpw, torch, h, stream, scans, steps, warm=3):
    """Batch resident in HBM -> (scans/s, ms per step, labels tensor) for `steps` timed steps on `stream`."""
    dev = torch.device("cuda", torch.cuda.current_device())
    off = np.zeros(len(scans) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in scans])
    d_pts = torch.from_numpy(np.concatenate(scans)).to(dev)
    d_lab = torch.empty(int(off[-1]), dtype=torch.uint8, device=dev)
    for _ in range(warm):
        h.segment_device(d_pts.data_ptr(), off, d_lab.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        h.segment_device(d_pts.data_ptr(), off, d_lab.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return len(scans) / (ms * 1e-3), ms, d_lab


def latency_stats(ts):
    ts = np.asarray(ts) * 1e3
    return {"p50": float(np.median(ts)), "p90": float(np.quantile(ts, 0.9)), "p99": float(np.quantile(ts, 0.99)), "calls": int(len(ts))}


def single_scan_latency(rpw, cfg, frames, device, solver_id, reps=200):
    """(i) synchronous rpw_segment on pinned host buffers (xyz stride 12 in, labels out), frames of varying size in
    rotation like a real stream; (ii) the C++ drop-in on pageable std::vector input returning the two clouds."""
    cap = max(len(a) for a in frames)
    h = rpw.Handle(cfg.to_c(), device, cap + cap // 4, 1)
    h.set_plane_solver(solver_id)
    pin = [rpw.capi.PinnedArray((len(a), 3), np.float32) for a in frames]
    lab = [rpw.capi.PinnedArray((len(a),), np.uint8) for a in frames]
    for p, a in zip(pin, frames):
        p.array[:] = a[:, :3]

    def call(k):
        rc = h.lib.rpw_segment(h._h, pin[k].ptr, len(frames[k]), 12, lab[k].ptr, None)
        if rc != 0:
            raise RuntimeError(h.lib.rpw_last_error(h._h).decode())

    for k in range(16):
        call(k % len(frames))
    ts = []
    for r in range(reps):
        t0 = time.perf_counter(); call(r % len(frames)); ts.append(time.perf_counter() - t0)
    out = {"rpw_segment_pinned_ms": latency_stats(ts), "graph": dict(zip(("launches", "captures"), h.scan_graph()))}
    h.close()
    exe = ROOT / "tools" / "_bin" / "dropin_latency"
    if exe.exists():
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
            for a in frames:
                f.write(struct.pack("<I", len(a)))
                f.write(np.ascontiguousarray(a[:, :3], np.float32).tobytes())
            path = f.name
        try:
            env = dict(os.environ, RPW_PLANE_SOLVER=str(solver_id), CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(device)))
            r = subprocess.run([str(exe), path, str(len(frames)), str(cfg.filtering_radius), str(reps)], capture_output=True, text=True, timeout=300, env=env)
            if r.returncode == 0:
                out["cpp_dropin_pageable_ms"] = json.loads(r.stdout.strip().splitlines()[-1])
            else:
                out["cpp_dropin_pageable_ms"] = {"error": r.stderr[-300:]}
        finally:
            os.unlink(path)
    return out


def e2e_rate(rpw, torch, cfg, scans, device, solver_id, n_handles, steps, world=1, dist=None, stride=12):
    """The public asynchronous C-ABI call on pinned host buffers, the batch cut into chunks rotating over a few handles so
    that H2D of one chunk, kernels of another and D2H of a third overlap; copies inside the timed region."""
    n_pts = [len(s) for s in scans]
    off = np.zeros(len(scans) + 1, np.int64)
    off[1:] = np.cumsum(n_pts)
    total = int(off[-1])
    words = stride // 4
    pin_in = rpw.capi.PinnedArray((total, words), np.float32)
    flat = np.concatenate(scans)
    pin_in.array[:, :3] = flat[:, :3]
    pin_out = rpw.capi.PinnedArray((total,), np.uint8)
    B = len(scans)
    n_chunks = max(1, min(n_handles, B))
    bounds = [round(i * B / n_chunks) for i in range(n_chunks + 1)]
    chunks = []
    for c in range(n_chunks):
        lo, hi = bounds[c], bounds[c + 1]
        if hi <= lo:
            continue
        hc = rpw.Handle(cfg.to_c(), device, int(off[hi] - off[lo]), hi - lo)
        hc.set_plane_solver(solver_id)
        chunks.append((hc, [pin_in.ptr + int(off[i]) * stride for i in range(lo, hi)], n_pts[lo:hi], [pin_out.ptr + int(off[i]) for i in range(lo, hi)]))

    def step():
        for hc, ip, ns, op in chunks:
            hc.wait()  # the previous step's labels of this chunk are complete in pinned host memory
            hc.segment_batch_async(ip, ns, stride, op)

    def drain():
        for hc, _, _, _ in chunks:
            hc.wait()

    for _ in range(3):
        step()
    drain()
    first = pin_out.array[: n_pts[0]].copy()
    launches0 = sum(hc.kernel_launches() for hc, _, _, _ in chunks)
    if dist is not None and world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    drain()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = sum(hc.kernel_launches() for hc, _, _, _ in chunks) - launches0
    for hc, _, _, _ in chunks:
        hc.close()
    return dt, first, total, len(chunks), launches


def main():
    # Libraries (NCCL's version banner, for one) print to stdout; the contract is ONE JSON line there.
    # Everything but that line goes to stderr: fd 1 is pointed at fd 2 and the line is written to the
    # saved descriptor at the end.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(1), "w", buffering=1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans", type=int, default=512, help="scans per step per GPU (BASELINE configs[2]: 4096 scans over 8 GPUs)")
    ap.add_argument("--solver", default="hybrid", choices=["eigen_qr", "closed_form", "hybrid", "reference"],
                    help="plane-normal solver: hybrid = the library default, eigen_qr = the reference's float QR throughout, "
                         "reference = QR + the reference's summation order (bit-identical labels; see include/rpw_b200.h)")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="handles the end-to-end arm ping-pongs over (copy/compute overlap)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU-seconds of work in the cpu_baseline sample (summed over threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shapes", action="store_true", help="skip the other named shapes (C1, C4, C5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, emit)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    rpw = importlib.import_module(PKG)
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    B = args.scans
    threads = max(1, (os.cpu_count() or 8) // max(1, min(world, 8)))
    seeds = range(1000 + rank * B, 1000 + rank * B + B)  # frame f of the stream -> its owner rank's block
    scans = gen_scans(rpw, seeds, threads)
    n_pts = [len(s) for s in scans]
    total = int(sum(n_pts))
    offsets = np.zeros(B + 1, np.uint64)
    offsets[1:] = np.cumsum(n_pts)

    h = rpw.Handle(cfg.to_c(), local_rank, total, B)
    solver_ids = {"eigen_qr": rpw.capi.SOLVER_EIGEN_QR, "closed_form": rpw.capi.SOLVER_CLOSED_FORM, "hybrid": rpw.capi.SOLVER_HYBRID,
                  "reference": rpw.capi.SOLVER_REFERENCE}
    solver_id = solver_ids[args.solver]
    h.set_plane_solver(solver_id)
    # a real (non-NULL) stream: the C-ABI reads NULL as "the handle's own stream", and CUDA events
    # only see the stream they are recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)

    # ---- device-resident arm ---------------------------------------------------------------
    host_f4 = np.concatenate(scans)  # (total, 4) float32
    d_pts = torch.from_numpy(host_f4).to(dev)
    d_labels = torch.empty(total, dtype=torch.uint8, device=dev)

    def step_resident():
        h.segment_device(d_pts.data_ptr(), offsets, d_labels.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_ranks(vals):
        """list of floats -> [world][len] on every rank"""
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world == 1:
            return [t.tolist()]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.tolist() for o in out]

    sampler = ClockSampler(local_rank)  # samples from the first warm-up step to the end of the e2e arm (all under load)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    h.profile_enable(True)
    launches0 = h.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = h.profile_read()
    h.profile_enable(False)
    launches = h.kernel_launches() - launches0
    ms_max = max_over_ranks(ms_total)
    value = world * B * args.steps / (ms_max * 1e-3)
    per_rank = gather_ranks([ms_total / args.steps] + [prof[k]["ms"] / max(1, prof[k]["launches"]) for k in ("bin", "offsets", "scatter", "fit")])

    # sanity: labels of the resident arm equal a host-path call on the first scan
    lab_dev = d_labels[: n_pts[0]].cpu().numpy()

    # ---- the same steps alternating over two handles on two streams: one batch's long-tail patches
    # overlap the next batch's bulk (reported beside the headline; `value` stays single-handle so that
    # the per-kernel times below add up to the step time) ----
    h2 = rpw.Handle(cfg.to_c(), local_rank, total, B)
    h2.set_plane_solver(solver_id)
    stream2 = torch.cuda.Stream(device=dev)
    h2.set_stream(stream2.cuda_stream)
    d_labels2 = torch.empty(total, dtype=torch.uint8, device=dev)
    pair = [(h, d_labels), (h2, d_labels2)]
    for r in range(4):
        pair[r % 2][0].segment_device(d_pts.data_ptr(), offsets, pair[r % 2][1].data_ptr())
    barrier()
    e0.record(stream)
    stream2.wait_event(e0)
    for r in range(args.steps):
        pair[r % 2][0].segment_device(d_pts.data_ptr(), offsets, pair[r % 2][1].data_ptr())
    ev2 = torch.cuda.Event()
    ev2.record(stream2)
    stream.wait_event(ev2)
    e1.record(stream)
    barrier()
    pipelined_value = world * B * args.steps / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
    assert torch.equal(d_labels, d_labels2)
    h2.close()
    del d_labels2

    # ---- the other solvers, same timed loop (reported beside the headline, not instead of it) ----
    def timed_with(sid, steps):
        h.set_plane_solver(sid)
        for _ in range(2):
            step_resident()
        barrier()
        e0.record(stream)
        for _ in range(steps):
            step_resident()
        e1.record(stream)
        barrier()
        return world * B * steps / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)

    other_values = {name: timed_with(sid, args.steps if name != "reference" else max(2, args.steps // 5))
                    for name, sid in solver_ids.items() if sid != solver_id}
    # label agreement of the timed solver with the reference-order mode (whose labels are the reference's, bit for bit:
    # tests/test_gpu_parity.py, tests/test_gpu_soak.py) and with the QR sequence on tree sums, over this rank's batch
    h.set_plane_solver(rpw.capi.SOLVER_REFERENCE)
    step_resident()
    torch.cuda.synchronize()
    labels_ref = d_labels.clone()
    h.set_plane_solver(rpw.capi.SOLVER_EIGEN_QR)
    step_resident()
    torch.cuda.synchronize()
    labels_qr = d_labels.clone()
    h.set_plane_solver(solver_id)
    step_resident()
    torch.cuda.synchronize()
    n_diff_vs_qr = int((labels_qr != d_labels).sum().item())
    diff_ref = (labels_ref != d_labels)
    n_diff_vs_ref = int(diff_ref.sum().item())
    worst_scan_vs_ref = 1.0
    if n_diff_vs_ref:
        per_scan = torch.zeros(B, dtype=torch.float64, device=dev)
        scan_of = torch.repeat_interleave(torch.arange(B, device=dev), torch.tensor(n_pts, device=dev))
        per_scan.index_add_(0, scan_of, diff_ref.double())
        worst_scan_vs_ref = float((1.0 - per_scan / torch.tensor(n_pts, dtype=torch.float64, device=dev)).min().item())
        del scan_of
    n_sample = min(8, B)  # kept for the cpu_baseline leg: the strict reference build checks these scans' labels in this run
    sample_pts = int(sum(n_pts[:n_sample]))
    sample_labels_ref_mode = labels_ref[:sample_pts].cpu().numpy()
    sample_labels_timed = d_labels[:sample_pts].cpu().numpy()
    del labels_qr, labels_ref, diff_ref
    parity_counts = gather_ranks([float(n_diff_vs_ref), float(total), worst_scan_vs_ref])

    # ---- end-to-end arm: pinned host xyz (12 B/pt, the reference's Point3D layout) -> labels ----
    e2e_steps = max(3, args.steps // 2)
    dt, first, _, n_handles, launches_e2e = e2e_rate(rpw, torch, cfg, scans, local_rank, solver_id, args.e2e_chunks, e2e_steps, world, dist)
    assert np.array_equal(first, lab_dev), "host-path and device-path labels differ"
    e2e_value = world * B * e2e_steps / max_over_ranks(dt)
    # the ceiling: what the ranks get from the host TOGETHER when they only copy (pinned host -> device 12 B/pt, labels back
    # 1 B/pt), every rank at the same time on its own GPU
    barrier()
    copy_bytes = min(total * 12, 512 << 20)
    reps = 6
    mix = rpw.capi.copy_probe(local_rank, copy_bytes, reps, both=True)  # the path's own mix: 12 B in + 1 B out per point, both engines
    barrier()
    h2d = rpw.capi.copy_probe(local_rank, copy_bytes, reps)
    barrier()
    h2d_wc = rpw.capi.copy_probe(local_rank, copy_bytes, reps, write_combined=True)
    barrier()
    d2h = rpw.capi.copy_probe(local_rank, max(1, copy_bytes // 12), reps, d2h=True)
    copy_rates = gather_ranks([mix, h2d, h2d_wc, d2h])
    clocks = sampler.stop()

    # ceiling: every rank moving 12 B per point in and 1 B out at the same time as every other rank, nothing else running
    e2e_ceiling = sum(r[0] * 1e9 / (POINTS_PER_SCAN * 12) for r in copy_rates)

    lat, shapes, pc2 = None, None, None
    if rank == 0:
        lat = {"C2-120k": single_scan_latency(rpw, cfg, scans[:8], local_rank, solver_id)}
        # ---- PointCloud2 ingest end to end: 32-byte records (x y z intensity ring time pad), xyz adjacent ----
        pc2_scans = scans[: min(B, 128)]
        dtp, _, _, _, _ = e2e_rate(rpw, torch, cfg, pc2_scans, local_rank, solver_id, args.e2e_chunks, 3, stride=32)
        pc2 = {"value": len(pc2_scans) * 3 / dtp, "unit": UNIT, "point_step": 32,
               "note": "rpw_segment_batch_async on whole 32-byte records (the generic strided read); rpw_segment_pc2 moves only the 12 xyz bytes per record when they are adjacent"}
        rec = np.zeros((len(scans[0]), 8), np.float32)
        rec[:, :3] = scans[0][:, :3]
        pin_rec = rpw.capi.PinnedArray(rec.shape, np.float32)
        pin_rec.array[:] = rec
        pin_lab = rpw.capi.PinnedArray((len(rec),), np.uint8)
        for packed in (1, 0):
            os.environ["RPW_PC2_PACK"] = str(packed)
            hq = rpw.Handle(cfg.to_c(), local_rank, POINTS_PER_SCAN + 4096, 1)
            ts = []
            for r in range(60):
                t0 = time.perf_counter()
                hq.lib.rpw_segment_pc2(hq._h, pin_rec.ptr, len(rec), 32, 0, 4, 8, pin_lab.ptr, None)
                ts.append(time.perf_counter() - t0)
            pc2["single_scan_ms_xyz_only_copy" if packed else "single_scan_ms_whole_records"] = latency_stats(ts[10:])
            assert np.array_equal(pin_lab.array, lab_dev)
            hq.close()
        os.environ.pop("RPW_PC2_PACK", None)

        # ---- multi-LiDAR frame with the fusion front end on the device: three sensors' clouds in their own
        # frames (a C4 frame cut into its three 120-degree sectors and rotated back), pageable host buffers in, labels out ----
        fused = None
        if not args.no_shapes:
            merged = rpw.synth.solidstate_merged(2000)[:, :3]
            az = np.degrees(np.arctan2(merged[:, 1], merged[:, 0]))
            yaws, clouds = [0.0, 120.0, -120.0], []
            for yaw in yaws:
                sel = merged[np.abs((az - yaw + 180.0) % 360.0 - 180.0) <= 60.0]
                a = np.radians(-yaw)
                rot = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]], np.float32)
                c = sel.copy()
                c[:, :2] = sel[:, :2] @ rot.T
                clouds.append(np.ascontiguousarray(c, np.float32))
            egos = [2.5, 2.5, 2.5]
            hf = rpw.Handle(rpw.PatchworkConfig().to_c(), local_rank, sum(map(len, clouds)) + 4096, 1)
            hf.set_plane_solver(solver_id)
            ts = []
            for r in range(60):
                t0 = time.perf_counter()
                lab_f = hf.segment_fused(clouds, yaws, egos)
                ts.append(time.perf_counter() - t0)
            hf.close()
            got = np.concatenate(lab_f)
            fused = {"sensors": 3, "points": int(sum(map(len, clouds))), "rpw_segment_fused_ms": latency_stats(ts[10:]),
                     "ground_fraction": float((got == 1).mean()), "ego_removed": int((got == 4).sum()),
                     "api": "rpw_segment_fused: rotation + ego removal + merge inside the binning kernel, pageable per-sensor buffers"}
            if not args.no_cpu_baseline:
                ref, kind, oracle_lib = load_reference()
                tr = []
                for r in range(5):
                    t0 = time.perf_counter()
                    out = ref.fuse(clouds, yaws, egos)
                    merged_cpu = out[0] if isinstance(out, tuple) else out
                    t1 = time.perf_counter()
                    ref.time_scan(oracle_lib.to_cfg(rpw.PatchworkConfig()), merged_cpu, 1)
                    tr.append((t1 - t0, time.perf_counter() - t1))
                fused["cpu_reference_ms"] = {"fuse": 1e3 * float(np.median([a for a, _ in tr])), "segment": 1e3 * float(np.median([b for _, b in tr])),
                                             "kind": kind, "cores": 1, "note": "LidarFusion::fuseLidarPointClouds + filterGroundPoints, one frame on one core"}

        # ---- the other named shapes (BASELINE configs[0], [3], [4]): resident batch, end to end, single-scan latency,
        # the reference's CPU build on the same host threads; parity for them is in tests/ ----
        if not args.no_shapes:
            shapes = {}
            peak, _ = measured_peaks()
            ncpu = os.cpu_count() or 1
            for name, pc, gen, base, nb in (
                    ("C1 reference test-suite cloud, 10k pts, defaults (BASELINE configs[0])", rpw.PatchworkConfig(), lambda s: rpw.synth.testsuite_cloud(int(s), 10000), 42, 1024),
                    ("C4 3x solid-state merged ~300k pts, banked track, R=150 (configs[3])", rpw.PatchworkConfig(), lambda s: rpw.synth.solidstate_merged(int(s)), 2000, 64),
                    ("C5 128-beam dense urban 262k pts, deep recursion, R=80 (configs[4])", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.dense_urban_scan(int(s)), 3000, 64)):
                sc = gen_scans(rpw, range(base, base + nb), threads, gen)
                pts = int(sum(len(a) for a in sc))
                hs = rpw.Handle(pc.to_c(), local_rank, pts, nb)
                hs.set_plane_solver(solver_id)
                hs.set_stream(stream.cuda_stream)
                rate, ms, _ = resident_rate(rpw, torch, hs, stream, sc, 6)
                hs.close()
                dts, _, _, _, _ = e2e_rate(rpw, torch, pc, sc, local_rank, solver_id, args.e2e_chunks, 3)
                entry = {"scans_per_step": nb, "points_per_scan_mean": pts / nb, "value": rate, "unit": "scans/s", "ms_per_step": ms,
                         "mpoints_per_sec": rate * pts / nb / 1e6, "hbm_fraction_whole_path": ALG_BYTES_PER_POINT * rate * pts / nb / 1e9 / peak,
                         "e2e": {"value": nb * 3 / dts, "unit": "scans/s", "h2d_bytes_per_step": pts * 12, "d2h_bytes_per_step": pts},
                         "single_scan_latency": single_scan_latency(rpw, pc, sc[:8], local_rank, solver_id, reps=100)}
                if not args.no_cpu_baseline:
                    v, kind, sample = cpu_reference_throughput(sc[:16], pc, max(4.0, args.cpu_seconds / 3), ncpu, what="this shape's seeds")
                    entry["cpu_baseline"] = {"value": v, "unit": "scans/s", "cores": ncpu, "kind": kind, "sample": sample}
                    v1, _, sample1 = cpu_reference_throughput(sc[:16], pc, 1.5, 1, what="this shape's seeds")
                    entry["cpu_baseline"]["one_core"] = {"value": v1, "unit": "scans/s", "sample": sample1}
                shapes[name] = entry

    if rank == 0:
        peak, peak_src = measured_peaks()
        fit = prof["fit"]
        fit_ms = fit["ms"] / max(1, fit["launches"])
        pts_per_launch = total * args.steps / max(1, fit["launches"])
        achieved = ALG_BYTES_PER_POINT * pts_per_launch / (fit_ms * 1e-3) / 1e9 if fit_ms > 0 else 0.0
        traffic, traffic_src, tj = None, None, {}
        tf = ROOT / "profiles" / "fit_traffic.json"
        if tf.exists():
            try:
                tj = json.loads(tf.read_text())
                traffic = tj["fit_phase_dram_bytes_per_point"] * pts_per_launch  # ncu, per launch
                traffic_src = tj.get("source")
            except Exception:
                traffic = None
        kernels, roof_k = {}, {}
        for k in ("bin", "offsets", "scatter", "fit"):
            ms = prof[k]["ms"] / max(1, prof[k]["launches"])
            kernels[k] = {"ms_per_launch": ms, "launches": prof[k]["launches"],
                          "share": prof[k]["ms"] / max(1e-9, sum(prof[q]["ms"] for q in ("bin", "offsets", "scatter", "fit")))}
        for k, bpp in PASS_BYTES.items():
            ms = kernels[k]["ms_per_launch"]
            ach = bpp * pts_per_launch / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            roof_k[k] = {"bound": "hbm", "pass_model_bytes_per_point": bpp, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "ms_per_launch": ms, "traffic": (tj.get(f"{k}_dram_bytes_per_point") or 0) * pts_per_launch or None}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(B, world),
            "mpoints_per_sec": value * POINTS_PER_SCAN / 1e6,
            "hbm_fraction_whole_path": ALG_BYTES_PER_POINT * (value / world) * POINTS_PER_SCAN / 1e9 / peak,
            "roofline": {"bound": "hbm", "kernel": "fit phase = rpw_fit_roots_kernel<64|128|256|512> (seven size classes on concurrent prioritised streams) + rpw_fit_levels_kernel, timed as one unit", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_point": ALG_BYTES_PER_POINT, "ms_per_launch": fit_ms,
                         "note": "HBM is not what bounds this kernel: it is instruction-issue bound (profiles/r02_fit_issue_bound.md)"},
            "roofline_kernels": roof_k,
            "kernels": kernels,
            "per_rank": {"columns": ["ms_per_step", "bin_ms", "offsets_ms", "scatter_ms", "fit_ms"], "rows": per_rank},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": total * 12, "d2h_bytes_per_step": total,
                    "api": f"rpw_segment_batch_async + rpw_wait (C-ABI) rolling over {n_handles} handles, pinned host xyz stride 12 in, labels out",
                    "ceiling_scans_per_sec": e2e_ceiling, "frac_of_ceiling": e2e_value / e2e_ceiling if e2e_ceiling else None,
                    "ceiling_note": "all ranks copying at the same time and nothing else (rpw_copy_probe): 12 B per point host->device with 1 B per point device->host on the other copy engine, summed over ranks",
                    "copy_gbs_per_rank": {"columns": ["h2d_with_d2h_mix", "h2d_alone", "h2d_alone_write_combined", "d2h_alone"], "rows": copy_rates},
                    "steps": e2e_steps, "gpu_launches": int(launches_e2e)},
            "e2e_pointcloud2": pc2,
            "fused_multi_lidar_frame": fused,
            "single_scan_latency": lat,
            "shapes": shapes,
            "pipelined_two_handles": {"value": pipelined_value, "unit": UNIT,
                                      "note": "same steps alternating over two handles / streams (frame-level pipelining by the caller)"},
            "solver": args.solver,
            "other_solvers": {k: {"value": v, "unit": UNIT} for k, v in other_values.items()},
            "parity": {"labels_differing_from_reference_order_mode": int(sum(r[0] for r in parity_counts)), "of": int(sum(r[1] for r in parity_counts)),
                       "worst_scan_agreement": min(r[2] for r in parity_counts),
                       "labels_differing_from_eigen_qr_rank0": n_diff_vs_qr,
                       "note": "all ranks' batches, timed solver against RPW_SOLVER_REFERENCE, whose labels are the reference's bit for bit (tests/test_gpu_parity.py, tests/test_gpu_soak.py)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, kind, sample = cpu_reference_throughput(scans[:16], cfg, args.cpu_seconds, os.cpu_count() or 1)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind, "sample": sample}
            v1, _, sample1 = cpu_reference_throughput(scans[:16], cfg, 3.0, 1)  # one core next to all cores
            out["cpu_baseline"]["one_core"] = {"value": v1, "unit": UNIT, "mpoints_per_sec": v1 * POINTS_PER_SCAN / 1e6, "sample": sample1}
            out["cpu_baseline"]["mpoints_per_sec"] = v * POINTS_PER_SCAN / 1e6
            # the same leg doubles as the checker: the reference's strict-IEEE build labels the first scans of the batch
            sys.path.insert(0, str(ROOT / "tests"))
            import oracle_lib
            strict = oracle_lib.try_reference("strict")
            if strict is not None:
                runs = [strict.run(oracle_lib.to_cfg(cfg), a) for a in scans[:n_sample]]
                ref_labels = np.concatenate([r["labels"] for r in runs])
                out["cpu_baseline"]["parity_sample"] = {
                    "scans": n_sample, "points": int(len(ref_labels)), "bit_equal_duplicate_points": int(sum(r["ambiguous"] for r in runs)),
                    "labels_differing_reference_order_mode_vs_libref_strict": int((ref_labels != sample_labels_ref_mode).sum()),
                    "labels_differing_timed_solver_vs_libref_strict": int((ref_labels != sample_labels_timed).sum())}
        emit(out)
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
