"""A wider run of tests/test_gpu_soak.py's comparison (same shapes, eight times the scans, seeds the gated test does not
use): keys and the three arithmetic modes against the CPU oracle, one summary line per shape.

    python tools/gpu_soak_wide.py [scale] > profiles/parity_soak_wide_r02.txt
"""
import importlib, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib

BAR = 0.999


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
    P = rpw.PatchworkConfig
    shapes = [
        ("C2", P(filtering_radius=80.0), lambda s: rpw.synth.spinning_scan(s), 50000, 64 * scale),
        ("C2nan", P(filtering_radius=80.0), lambda s: rpw.synth.spinning_scan(s, nan_per_million=5000), 60000, 8 * scale),
        ("C5", P(filtering_radius=80.0), lambda s: rpw.synth.dense_urban_scan(s), 31000, 48 * scale),
        ("C4", P(), lambda s: rpw.synth.solidstate_merged(s), 21000, 24 * scale),
        ("C1", P(), lambda s: rpw.synth.testsuite_cloud(s, 10000), 1000, 32 * scale),
        ("C1pct", P(adaptive_seed_height=False, num_sectors=37), lambda s: rpw.synth.testsuite_cloud(s, 10000), 3000, 16 * scale),
    ]
    oracle = oracle_lib.Oracle()
    tot = np.zeros(6, np.int64)
    for name, cfg, gen, seed0, count in shapes:
        ocfg = oracle_lib.to_cfg(cfg)
        rows = []
        h = None
        for lo in range(0, count, 64):  # in slices, so that the host never holds more than 64 scans and their oracle results
            seeds = range(seed0 + lo, seed0 + min(count, lo + 64))
            with ThreadPoolExecutor(16) as ex:
                scans = list(ex.map(gen, seeds))
                want = list(ex.map(lambda a: oracle.run(ocfg, a), scans))
            if h is None:
                h = rpw.Handle(cfg.to_c(), 0, 400000, 1)
            for seed, a, o in zip(seeds, scans, want):
                h.set_plane_solver(rpw.capi.SOLVER_HYBRID); h.set_exact_replay(-1)
                fast = h.segment(a)
                keys_bad = int((h.debug_keys(len(a)) != o["keys"]).sum())
                h.set_exact_replay(8)
                replay = h.segment(a)
                h.set_exact_replay(-1); h.set_plane_solver(rpw.capi.SOLVER_REFERENCE)
                exact = h.segment(a)
                rows.append((seed, len(a), keys_bad, int((fast != o["labels"]).sum()), int((replay != o["labels"]).sum()), int((exact != o["labels"]).sum())))
        h.close()
        n = sum(r[1] for r in rows)
        below_fast = [(r[0], round(1 - r[3] / r[1], 5)) for r in rows if r[3] > (1 - BAR) * r[1]]
        below_replay = [(r[0], round(1 - r[4] / r[1], 5)) for r in rows if r[4] > (1 - BAR) * r[1]]
        print(f"{name}: scans {len(rows)} points {n} | key mismatches {sum(r[2] for r in rows)} | labels differing from the oracle: "
              f"reference-order mode {sum(r[5] for r in rows)}, default solver {sum(r[3] for r in rows)} ({100 * (1 - sum(r[3] for r in rows) / n):.5f} % agree), "
              f"replay(8) {sum(r[4] for r in rows)} | scans below 99.9 %: default {below_fast or 'none'}, replay(8) {below_replay or 'none'}", flush=True)
        tot += np.array([len(rows), n, sum(r[2] for r in rows), sum(r[3] for r in rows), sum(r[4] for r in rows), sum(r[5] for r in rows)])
    print(f"TOTAL: scans {tot[0]} points {tot[1]} | key mismatches {tot[2]} | reference-order mode {tot[5]} labels differing | default solver {tot[3]} "
          f"({100 * (1 - tot[3] / tot[1]):.5f} % agree) | replay(8) {tot[4]}")


if __name__ == "__main__":
    main()
