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
timed region.  `roofline` is for the dominant kernel (fit), timed live with CUDA events on the
launching stream.  `cpu_baseline` / `--impl reference` time the reference's own CPU implementation
(oracle/_ref/libref_fast.so, the reference's translation units built with its library flags) on the
box's host cores.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
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
METRIC = "scans_per_sec"
UNIT = "scans/s (120k-point 64-beam scans)"


def gen_scans(rpw, seeds, threads):
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(lambda s: rpw.synth.spinning_scan(int(s)), seeds))


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


def cpu_reference_throughput(scans, cfg, seconds_budget, threads):
    """Reference CPU implementation on the host cores: frame-parallel, one RecursivePatchwork per call
    (the class is stateless, RP/include/recursive_patchwork.hpp:70).  Returns scans/s and a description."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib
    ref = oracle_lib.try_reference("fast")
    kind = "reference"
    if ref is None:
        ref = oracle_lib.Oracle()  # the C restatement ("port") when oracle/_ref was not built
        kind = "port"
    ccfg = oracle_lib.to_cfg(cfg)
    t_one = ref.time_scan(ccfg, scans[0], 1)  # warm-up + estimate
    n_jobs = max(threads, min(4000, int(seconds_budget / max(t_one, 1e-4))))  # ~seconds_budget CPU-seconds of work in total
    jobs = [scans[i % len(scans)] for i in range(n_jobs)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda a: ref.time_scan(ccfg, a, 1), jobs))
    dt = time.perf_counter() - t0
    return n_jobs / dt, kind, f"{n_jobs} scans ({len(scans)} distinct, C2 seeds) on {threads} threads, {dt:.1f} s wall, libref_{'fast' if kind == 'reference' else 'oracle'}"


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rpw = importlib.import_module(PKG)
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    threads = os.cpu_count() or 1
    scans = gen_scans(rpw, range(1000, 1000 + 16), threads)
    per_step = max(threads * 16, 64)  # large enough that the thread pool's start-up does not show
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib
    ref = oracle_lib.try_reference("fast")
    kind = "reference" if ref is not None else "port"
    if ref is None:
        ref = oracle_lib.Oracle()
    ccfg = oracle_lib.to_cfg(cfg)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        def step():
            list(ex.map(lambda i: ref.time_scan(ccfg, scans[i % len(scans)], 1), range(per_step)))

        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    sample = f"{per_step} scans per step ({len(scans)} distinct C2 scans) on {threads} host threads"
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 KITTI-like 64-beam 120k-point scans, R=80 m, default zone model", "scans_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mpoints_per_sec": val * POINTS_PER_SCAN / 1e6,
    })


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
    ap.add_argument("--solver", default="hybrid", choices=["eigen_qr", "closed_form", "hybrid"],
                    help="plane-normal solver: hybrid = the library default, eigen_qr = the reference's float QR throughout (see include/rpw_b200.h)")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="handles the end-to-end arm ping-pongs over (copy/compute overlap)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU-seconds of work in the cpu_baseline sample (summed over threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    solver_id = {"eigen_qr": rpw.capi.SOLVER_EIGEN_QR, "closed_form": rpw.capi.SOLVER_CLOSED_FORM, "hybrid": rpw.capi.SOLVER_HYBRID}[args.solver]
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

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

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
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pipelined_value = world * B * args.steps / (float(t.item()) * 1e-3)
    assert torch.equal(d_labels, d_labels2)
    h2.close()

    # ---- the other solvers, same timed loop (reported beside the headline, not instead of it) ----
    def timed_with(sid):
        h.set_plane_solver(sid)
        for _ in range(3):
            step_resident()
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
        e1.record(stream)
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * B * args.steps / (float(tt.item()) * 1e-3)

    names = {rpw.capi.SOLVER_EIGEN_QR: "eigen_qr", rpw.capi.SOLVER_CLOSED_FORM: "closed_form", rpw.capi.SOLVER_HYBRID: "hybrid"}
    other_values = {names[sid]: timed_with(sid) for sid in names if sid != solver_id}
    # label agreement of the timed solver with the reference's own QR sequence over this rank's batch
    h.set_plane_solver(rpw.capi.SOLVER_EIGEN_QR)
    step_resident()
    torch.cuda.synchronize()
    labels_qr = d_labels.clone()
    h.set_plane_solver(solver_id)
    step_resident()
    torch.cuda.synchronize()
    n_diff_vs_qr = int((labels_qr != d_labels).sum().item())
    del labels_qr

    # ---- end-to-end arm: pinned host xyz (12 B/pt, the reference's Point3D layout) -> labels ----
    # The public C-ABI call a user makes (rpw_segment_batch_async + rpw_wait) on host buffers; the
    # batch is cut into chunks that rotate over a few handles so that the H2D copy of one chunk, the
    # kernels of another and the D2H copy of a third overlap.  A handle is waited for (its labels are
    # on the host) right before it is given its chunk of the NEXT step, so the copy engine never idles
    # between steps; the timed region ends when every label of every step has arrived.
    pin_in = rpw.capi.PinnedArray((total, 3), np.float32)
    pin_in.array[:] = host_f4[:, :3]
    pin_out = rpw.capi.PinnedArray((total,), np.uint8)
    n_chunks = max(1, min(args.e2e_chunks, B))
    bounds = [round(i * B / n_chunks) for i in range(n_chunks + 1)]
    chunks = []
    for c in range(n_chunks):
        lo, hi = bounds[c], bounds[c + 1]
        if hi <= lo:
            continue
        hc = rpw.Handle(cfg.to_c(), local_rank, int(offsets[hi] - offsets[lo]), hi - lo)
        hc.set_plane_solver(solver_id)
        chunks.append((hc, [pin_in.ptr + int(offsets[i]) * 12 for i in range(lo, hi)], n_pts[lo:hi],
                       [pin_out.ptr + int(offsets[i]) for i in range(lo, hi)]))

    def step_e2e():
        for hc, ip, ns, op in chunks:
            hc.wait()  # the previous step's labels of this chunk are complete in pinned host memory
            hc.segment_batch_async(ip, ns, 12, op)

    def drain_e2e():
        for hc, _, _, _ in chunks:
            hc.wait()

    for _ in range(3):
        step_e2e()
    drain_e2e()
    assert np.array_equal(pin_out.array[: n_pts[0]], lab_dev), "host-path and device-path labels differ"
    e2e_steps = max(3, args.steps // 2)
    launches_e2e0 = sum(hc.kernel_launches() for hc, _, _, _ in chunks)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    drain_e2e()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t.item())
    for hc, _, _, _ in chunks:
        hc.close()
    clocks = sampler.stop()

    # ---- single-scan latency: the reference's own use (one scan per ROS2 callback), synchronous
    # rpw_segment on pinned host buffers, H2D + 11 launches + D2H per call ----
    lat_ms, other_shapes = None, None
    if rank == 0:
        h1 = rpw.Handle(cfg.to_c(), local_rank, POINTS_PER_SCAN + 4096, 1)
        h1.set_plane_solver(solver_id)
        lat = []
        for i in range(min(B, 48)):
            a = pin_in.array[int(offsets[i]):int(offsets[i + 1])]
            o = pin_out.array[int(offsets[i]):int(offsets[i + 1])]
            t0 = time.perf_counter()
            h1.lib.rpw_segment(h1._h, a.ctypes.data, len(a), 12, o.ctypes.data, None)
            lat.append(time.perf_counter() - t0)
        lat_ms = {"median": 1e3 * float(np.median(lat[8:])), "p90": 1e3 * float(np.quantile(lat[8:], 0.9)), "scans": len(lat) - 8}
        h1.close()
        # the other named shapes, one scan at a time (BASELINE configs[3] and [4]); parity for them is in tests/
        other_shapes = {}
        for name, pc, cloud in (("C4 3x solid-state merged ~300k pts, banked track, R=150", rpw.PatchworkConfig(), rpw.synth.solidstate_merged(2000)),
                                ("C5 128-beam dense urban 262k pts, deep recursion, R=80", rpw.PatchworkConfig(filtering_radius=80.0), rpw.synth.dense_urban_scan(3000))):
            hs = rpw.Handle(pc.to_c(), local_rank, len(cloud) + 4096, 1)
            hs.set_plane_solver(solver_id)
            pi = rpw.capi.PinnedArray((len(cloud), 3), np.float32)
            pi.array[:] = cloud[:, :3]
            po = rpw.capi.PinnedArray((len(cloud),), np.uint8)
            ts = []
            for _ in range(12):
                t0 = time.perf_counter()
                hs.lib.rpw_segment(hs._h, pi.ptr, len(cloud), 12, po.ptr, None)
                ts.append(time.perf_counter() - t0)
            med = float(np.median(ts[2:]))
            other_shapes[name] = {"points": int(len(cloud)), "ms_per_scan": 1e3 * med, "scans_per_sec": 1.0 / med, "mpoints_per_sec": len(cloud) / med / 1e6,
                                  "ground_points": int((po.array == 1).sum())}
            hs.close()

    if rank == 0:
        peak, peak_src = measured_peaks()
        fit = prof["fit"]
        fit_ms = fit["ms"] / max(1, fit["launches"])
        pts_per_launch = total * args.steps / max(1, fit["launches"])
        achieved = ALG_BYTES_PER_POINT * pts_per_launch / (fit_ms * 1e-3) / 1e9 if fit_ms > 0 else 0.0
        traffic = None
        tf = ROOT / "profiles" / "fit_traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text())["fit_phase_dram_bytes_per_point"] * pts_per_launch  # ncu, per launch
            except Exception:
                traffic = None
        kernels = {}
        for k in ("bin", "offsets", "scatter", "fit"):
            ms = prof[k]["ms"] / max(1, prof[k]["launches"])
            kernels[k] = {"ms_per_launch": ms, "launches": prof[k]["launches"],
                          "share": prof[k]["ms"] / max(1e-9, sum(prof[q]["ms"] for q in ("bin", "offsets", "scatter", "fit")))}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2 KITTI-like 64-beam 120k-point scans (BASELINE configs[1]) streamed as configs[2] batches, R=80 m, default zone model, seeds 1000+",
                       "scans_per_step_per_gpu": B, "points_per_scan": POINTS_PER_SCAN, "parallelism": f"frames x{world}",
                       "l2": f"batch is {total * 16 / 1e6:.0f} MB of float4 input per GPU > 126 MB L2 (no flush needed)"},
            "mpoints_per_sec": value * POINTS_PER_SCAN / 1e6,
            "hbm_fraction_whole_path": ALG_BYTES_PER_POINT * (value / world) * POINTS_PER_SCAN / 1e9 / peak,
            "roofline": {"bound": "hbm", "kernel": "fit phase = rpw_fit_roots_kernel<64|128|256|512> (seven size classes on concurrent prioritised streams) + rpw_fit_levels_kernel, timed as one unit", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_point": ALG_BYTES_PER_POINT, "ms_per_launch": fit_ms},
            "kernels": kernels,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": total * 12, "d2h_bytes_per_step": total,
                    "api": f"rpw_segment_batch_async + rpw_wait (C-ABI) rolling over {len(chunks)} handles, pinned host xyz stride 12 in, labels out",
                    "pcie_note": "host->device copies alone run at 54.4 GB/s on this box (tools/gpu_pcie.py): 37.8 k scans/s of 1.44 MB is the ceiling",
                    "steps": e2e_steps},
            "single_scan_latency_ms": lat_ms,
            "other_shapes_single_scan": other_shapes,
            "pipelined_two_handles": {"value": pipelined_value, "unit": UNIT,
                                      "note": "same steps alternating over two handles / streams (frame-level pipelining by the caller)"},
            "solver": args.solver,
            "other_solvers": {k: {"value": v, "unit": UNIT} for k, v in other_values.items()},
            "labels_differing_from_eigen_qr": {"count": n_diff_vs_qr, "of": total, "note": "this rank's batch, timed solver vs the reference's float QR sequence"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, kind, sample = cpu_reference_throughput(scans[:16], cfg, args.cpu_seconds, os.cpu_count() or 1)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind, "sample": sample}
        emit(out)
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
