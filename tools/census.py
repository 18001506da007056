"""Census of the named synthetic shapes from the CPU oracle: points, zone/binned counts,
non-empty patches, fitPlaneAndSplit nodes per depth, plane-fit iterations, splits by cause.
`python tests/census.py > profiles/census_r01.md` (CPU only)."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib  # noqa: E402

rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
PC, S = rpw.PatchworkConfig, rpw.synth
orc = oracle_lib.Oracle()
shapes = [
    ("C1 test-suite cloud 3,000 pts, seed 42, defaults", PC(), S.testsuite_cloud(42, 3000)),
    ("C1 test-suite cloud 5,000 pts, seed 43, R=50 S=8 max_iter=50", PC(filtering_radius=50.0, num_sectors=8, max_iter=50), S.testsuite_cloud(43, 5000)),
    ("C1 test-suite cloud 10,000 pts, seed 42, defaults", PC(), S.testsuite_cloud(42, 10000)),
    ("C2 64-beam spinning scan, seed 1000, R=80", PC(filtering_radius=80.0), S.spinning_scan(1000)),
    ("C2 64-beam spinning scan, seed 1001, R=80", PC(filtering_radius=80.0), S.spinning_scan(1001)),
    ("C4 3x solid-state merged, banked track, seed 2000, defaults (R=150)", PC(), S.solidstate_merged(2000)),
    ("C5 128-beam dense urban two-layer clutter, seed 3000, R=80", PC(filtering_radius=80.0), S.dense_urban_scan(3000)),
    ("C5 128-beam dense urban two-layer clutter, seed 3001, R=80", PC(filtering_radius=80.0), S.dense_urban_scan(3001)),
]
print("# Census of the synthetic shapes (CPU oracle)\n")
print("| shape | points | in zone | binned | patches | nodes | nodes per depth | splits (collapse route) | fit iterations | point-iterations / binned point | patches that split | patches reaching depth >= 4 | ground |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, cfg, pts in shapes:
    r = orc.run(cfg, pts, want_nodes=True)
    st, nd = r["stats"], r["nodes"]
    roots = np.unique(nd["root"])
    split = sum(1 for rt in roots if (nd[nd["root"] == rt]["outcome"] == 5).any())
    deep = sum(1 for rt in roots if nd[nd["root"] == rt]["depth"].max() >= 4)
    print(f"| {name} | {len(pts)} | {st['n_zone']} | {st['n_binned']} | {st['n_root_patches']} | {st['n_nodes']} | {np.bincount(nd['depth']).tolist()} | "
          f"{st['n_splits']} ({st['n_splits_collapse']}) | {st['n_pca_iters']} | {st['n_point_iters'] / max(1, st['n_binned']):.2f} | {split} | {deep} | {st['n_ground']} |")
