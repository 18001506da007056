"""How far do two builds of the UNMODIFIED reference agree with each other on the soak scans?  (strict IEEE flags
vs the flags its CMake ships, -O3 -ffast-math; oracle/_ref, so this runs only where /root/reference was compiled.)
Context for the GPU's own deviations on bistable / non-converging patches (profiles/parity_soak_r01.md)."""
import importlib, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import oracle_lib
strict, fast = oracle_lib.Reference("strict"), oracle_lib.Reference("fast")
SHAPES = [
    ("C2 64-beam 120k, R=80", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.spinning_scan(s, nan_per_million=0), range(5000, 5192)),
    ("C5 128-beam 262k two-layer clutter, R=80", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.spinning_scan(s, 128, 2048, 1, 0), range(3100, 3148)),
    ("C4 3 x solid-state 300k, R=150", rpw.PatchworkConfig(), lambda s: rpw.synth.solidstate_merged(s), range(2100, 2124)),
]
print("| shape | scans | labels differing between the reference's two builds | scans below 99.9 % | worst scan agreement |")
print("|---|---|---|---|---|")
with ThreadPoolExecutor(8) as ex:
    for name, cfg, gen, seeds in SHAPES:
        ocfg = oracle_lib.to_cfg(cfg)
        scans = list(ex.map(gen, seeds))
        a = list(ex.map(lambda p: strict.run(ocfg, p)["labels"], scans))
        b = list(ex.map(lambda p: fast.run(ocfg, p)["labels"], scans))
        d = [int((x != y).sum()) for x, y in zip(a, b)]
        print(f"| {name} | {len(scans)} | {sum(d)} of {sum(map(len, scans))} | {sum(1 for k, p in zip(d, scans) if k > 0.001 * len(p))} | {min(1 - k / len(p) for k, p in zip(d, scans)):.6f} |", flush=True)
