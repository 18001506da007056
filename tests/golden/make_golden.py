"""Generates tests/golden/*.npz from the reference ITSELF (oracle/_ref/libref_strict.so = the
reference's translation units compiled unmodified, see oracle/Makefile).  Run in the build
container, where /root/reference exists:

    python tests/golden/make_golden.py

Each fixture holds the input cloud (float32 xyz), the config, and the reference's outputs: the
per-input-point labels reconstructed from its two returned clouds, and the cloud sizes.  The
reference's own tests pin no numbers, so these are the golden vectors of this repo."""
import importlib
import sys
from dataclasses import asdict
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib  # noqa: E402

rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
PC, S = rpw.PatchworkConfig, rpw.synth


def cases():
    rng = np.random.default_rng(7)
    few = rng.uniform(-20, 20, (2, 3)).astype(np.float32)
    mixed = S.testsuite_cloud(45, 2000)[:, :3].copy()
    mixed[::97, 0] = np.nan
    mixed[5::131, 2] = np.inf
    mixed[::50, :2] *= 10.0  # beyond the radius
    dup = S.testsuite_cloud(46, 1500)[:, :3].copy()
    dup[:, 0] = np.round(dup[:, 0])  # heavy coordinate duplication
    dup[:, 1] = np.round(dup[:, 1] * 2) / 2
    return [
        ("c1_3000_s42", PC(), S.testsuite_cloud(42, 3000)[:, :3]),
        ("c1_5000_s43_r50_s8", PC(filtering_radius=50.0, num_sectors=8, max_iter=50), S.testsuite_cloud(43, 5000)[:, :3]),
        ("c1_10000_s42_splits", PC(), S.testsuite_cloud(42, 10000)[:, :3]),
        ("c1_10000_s42_nonadaptive", PC(adaptive_seed_height=False), S.testsuite_cloud(42, 10000)[:, :3]),
        ("c2_small_64x220", PC(filtering_radius=80.0), S.spinning_scan(1000, 64, 220)[:, :3]),
        ("c5_small_128x256", PC(filtering_radius=80.0), S.spinning_scan(3000, 128, 256, 1)[:, :3]),
        ("edge_two_points", PC(), few),
        ("edge_nan_inf_beyond", PC(filtering_radius=40.0), mixed),
        ("edge_duplicates", PC(filtering_radius=60.0), dup),
    ]


def main():
    ref = oracle_lib.Reference("strict")
    for name, cfg, pts in cases():
        pts = np.ascontiguousarray(pts, np.float32)
        r = ref.run(cfg, pts)
        np.savez_compressed(HERE / f"{name}.npz", points=pts, labels=r["labels"], n_ground=len(r["ground"]),
                            n_non_ground=len(r["non_ground"]), ambiguous=r["ambiguous"],
                            config=np.array([list(asdict(cfg).values())], dtype=np.float64), config_keys=np.array(list(asdict(cfg).keys())))
        print(f"{name}: n={len(pts)} ground={len(r['ground'])} non_ground={len(r['non_ground'])} ambiguous={r['ambiguous']}")


if __name__ == "__main__":
    main()
