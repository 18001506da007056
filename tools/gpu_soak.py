"""Parity soak: many scans of every shape through the C-ABI, both solvers, against the CPU oracle.
Writes a markdown table to stdout (kept as profiles/parity_soak_r01.md).  usage: gpu_soak.py [scale]"""
import importlib, sys, time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import oracle_lib
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
oracle = oracle_lib.Oracle()
SHAPES = [
    ("C2 64-beam 120k, R=80", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.spinning_scan(s), range(5000, 5000 + int(192 * scale))),
    ("C2 with 0.5 % non-finite points", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.spinning_scan(s, nan_per_million=5000), range(6000, 6000 + int(32 * scale))),
    ("C5 128-beam 262k two-layer clutter, R=80", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.dense_urban_scan(s), range(3100, 3100 + int(48 * scale))),
    ("C4 3 x solid-state 300k, R=150", rpw.PatchworkConfig(), lambda s: rpw.synth.solidstate_merged(s), range(2100, 2100 + int(24 * scale))),
    ("C1 test-suite cloud 10k, defaults", rpw.PatchworkConfig(), lambda s: rpw.synth.testsuite_cloud(s, 10000), range(100, 100 + int(128 * scale))),
    ("C1 10k, 10th-percentile seeds (adaptive off), 37 sectors", rpw.PatchworkConfig(adaptive_seed_height=False, num_sectors=37), lambda s: rpw.synth.testsuite_cloud(s, 10000), range(300, 300 + int(64 * scale))),
]
print("# Parity soak, round 1 (tools/gpu_soak.py): GPU labels and keys against the CPU oracle\n")
print("| shape | scans | points | key mismatches | labels differing, hybrid (default) | labels differing, eigen_qr | worst scan agreement (hybrid / eigen_qr) |")
print("|---|---|---|---|---|---|---|")
tot = dict(n=0, k=0, h=0, q=0, scans=0)
with ThreadPoolExecutor(16) as ex:
    for name, cfg, gen, seeds in SHAPES:
        scans = list(ex.map(gen, seeds))
        ocfg = oracle_lib.to_cfg(cfg)
        want = list(ex.map(lambda a: oracle.run(ocfg, a), scans))
        h = rpw.Handle(cfg.to_c(), 0, max(len(a) for a in scans) + 4096, 1)
        n = kbad = 0
        bad = {0: 0, 2: 0}; worst = {0: 1.0, 2: 1.0}
        for a, o in zip(scans, want):
            n += len(a)
            for sid in (2, 0):
                h.set_plane_solver(sid)
                lab = h.segment(a)
                d = int((lab != o["labels"]).sum())
                bad[sid] += d
                worst[sid] = min(worst[sid], 1.0 - d / max(1, len(a)))
            kbad += int((h.debug_keys(len(a)) != o["keys"]).sum())
        h.close()
        print(f"| {name} | {len(scans)} | {n} | {kbad} | {bad[2]} | {bad[0]} | {worst[2]:.6f} / {worst[0]:.6f} |", flush=True)
        tot["n"] += n; tot["k"] += kbad; tot["h"] += bad[2]; tot["q"] += bad[0]; tot["scans"] += len(scans)
print(f"| **total** | {tot['scans']} | {tot['n']} | {tot['k']} | {tot['h']} | {tot['q']} | |")
print(f"\nAgreement overall: hybrid {1 - tot['h'] / tot['n']:.8f}, eigen_qr {1 - tot['q'] / tot['n']:.8f} (bar: 0.999 per scan).")
