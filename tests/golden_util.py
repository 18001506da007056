"""Loads tests/golden/*.npz (made by tests/golden/make_golden.py from the reference itself)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
INT_KEYS = ("num_sectors", "max_iter", "max_split_depth")


def names():
    return sorted(p.stem for p in GOLDEN.glob("*.npz"))


def load(name, PatchworkConfig):
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    kw = {}
    for k, v in zip(z["config_keys"], z["config"][0]):
        k = str(k)
        kw[k] = int(v) if k in INT_KEYS else (bool(v) if k == "adaptive_seed_height" else float(np.float32(v)))
    return dict(points=z["points"], labels=z["labels"], n_ground=int(z["n_ground"]), n_non_ground=int(z["n_non_ground"]),
                cfg=PatchworkConfig(**kw))
