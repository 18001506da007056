"""Single-scan latency through rpw_segment (pinned host buffers in, labels out) for the named shapes, with the
per-kernel device times of the same calls.  usage: gpu_latency.py [reps]"""
import importlib, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
SHAPES = [("C1-10k", rpw.PatchworkConfig(), lambda s: rpw.synth.testsuite_cloud(s, 10000), 42),
          ("C2-120k", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.spinning_scan(s), 1000),
          ("C4-300k", rpw.PatchworkConfig(), lambda s: rpw.synth.solidstate_merged(s), 2000),
          ("C5-262k", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.dense_urban_scan(s), 3000)]
for name, cfg, gen, seed in SHAPES:
    scans = [np.ascontiguousarray(gen(seed + k)[:, :3]) for k in range(8)]   # varying point counts, like a real stream
    cap = max(len(a) for a in scans)
    h = rpw.Handle(cfg.to_c(), 0, cap + cap // 4, 1)
    pin = [rpw.capi.PinnedArray((len(a), 3), np.float32) for a in scans]
    lab = [rpw.capi.PinnedArray((len(a),), np.uint8) for a in scans]
    for p, a in zip(pin, scans): p.array[:] = a
    def call(k):
        h._check(h.lib.rpw_segment(h._h, pin[k].ptr, len(scans[k]), 12, lab[k].ptr, None))
    for k in range(16): call(k % 8)
    t = []
    for r in range(reps):
        k = r % 8
        t0 = time.perf_counter(); call(k); t.append((time.perf_counter() - t0) * 1e3)
    t = np.array(t)
    h.profile_enable(True)
    for r in range(40): call(r % 8)
    p = h.profile_read()
    ker = " ".join(f"{k} {p[k]['ms'] / max(1, p[k]['launches']) * 1e3:.1f}us" for k in ("bin", "offsets", "scatter", "fit"))
    print(f"{name:8s} n~{cap:7d}  p50 {np.median(t):.3f} ms  p90 {np.quantile(t, .9):.3f}  p99 {np.quantile(t, .99):.3f}  min {t.min():.3f}  "
          f"-> {1e3 / np.median(t):.0f} scans/s   kernels: {ker}", flush=True)
    h.close()
