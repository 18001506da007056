"""Diagnostic run on a GPU box: every named shape through the CUDA path next to the oracle, with
the parity report printed rather than asserted.  `python tools/gpu_diag.py [quick]`."""
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib  # noqa: E402
import parity  # noqa: E402

rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    orc = oracle_lib.Oracle()
    S = rpw.synth
    PC = rpw.PatchworkConfig
    cases = [
        ("C1-3k", PC(), S.testsuite_cloud(42, 3000)),
        ("C1-5k", PC(filtering_radius=50.0, num_sectors=8, max_iter=50), S.testsuite_cloud(43, 5000)),
        ("C1-10k", PC(), S.testsuite_cloud(42, 10000)),
        ("C1-10k-nonadaptive", PC(adaptive_seed_height=False), S.testsuite_cloud(42, 10000)),
        ("C2", PC(filtering_radius=80.0), S.spinning_scan(1000)),
    ]
    if not quick:
        cases += [
            ("C4", PC(), S.solidstate_merged(2000)),
            ("C5", PC(filtering_radius=80.0), S.dense_urban_scan(3000)),
            ("C5b", PC(filtering_radius=80.0), S.dense_urban_scan(3001)),
        ]
    h = rpw.Handle(None, 0, 1 << 20, 4)
    h.enable_nodes(True)
    for name, cfg, pts in cases:
        h.set_config(cfg.to_c())
        t0 = time.time()
        labels, st = h.segment(pts, want_stats=True)
        t1 = time.time()
        keys = h.debug_keys(len(pts))
        nodes = h.debug_nodes()
        o = orc.run(cfg, pts, want_nodes=True)
        rep = parity.compare_scan(labels, keys, o)
        nrep = parity.compare_nodes(nodes, o["nodes"])
        print(f"== {name}: n={len(pts)} gpu_wall_ms={1e3 * (t1 - t0):.2f} levels={st.n_levels} nodes={st.n_nodes}")
        print("   scan :", json.dumps(rep))
        print("   nodes:", json.dumps({k: (v if not isinstance(v, tuple) else list(map(str, v))) for k, v in nrep.items()}))
        print("   oracle stats:", {k: o["stats"][k] for k in ("n_nodes", "n_leaves", "n_splits", "max_depth", "n_pca_iters")})
        # 12-byte stride must give the same answer
        labels12 = h.segment(np.ascontiguousarray(pts[:, :3]))
        print("   stride12 == stride16:", bool(np.array_equal(labels12, labels)))
    # batch == singles
    cfg = PC(filtering_radius=80.0)
    h.set_config(cfg.to_c())
    scans = [S.spinning_scan(1000 + i) for i in range(4)]
    singles = [h.segment(s) for s in scans]
    batch = h.segment_batch(scans)
    print("batch == singles:", all(np.array_equal(a, b) for a, b in zip(singles, batch)))
    # repeatability
    again = h.segment_batch(scans)
    print("repeatable:", all(np.array_equal(a, b) for a, b in zip(again, batch)))
    # device math
    rng = np.random.default_rng(0)
    y = rng.uniform(-80, 80, 200000).astype(np.float32)
    x = rng.uniform(-80, 80, 200000).astype(np.float32)
    dev = h.debug_atan2(y, x)
    import ctypes as C
    libm = C.CDLL("libm.so.6")
    libm.atan2f.restype = C.c_float
    libm.atan2f.argtypes = [C.c_float, C.c_float]
    host = np.array([libm.atan2f(float(a), float(b)) for a, b in zip(y[:50000], x[:50000])], np.float32)
    print("atan2 device vs libm mismatches (50k):", int((dev[:50000].view(np.uint32) != host.view(np.uint32)).sum()))
    A = rng.normal(size=(20000, 3, 3)).astype(np.float32)
    A = (A @ A.transpose(0, 2, 1)).astype(np.float32)
    A[:, 2, 2] *= 1e-4
    A = ((A + A.transpose(0, 2, 1)) * 0.5).astype(np.float32)
    ev, vec = h.debug_eig3(A)
    oev, ovec = orc.eig3(A)
    print("eig3 device vs oracle: eval bit mismatches", int((ev.view(np.uint32) != oev.view(np.uint32)).sum()),
          "evec bit mismatches", int((vec.view(np.uint32) != ovec.view(np.uint32)).sum()))
    h.close()


if __name__ == "__main__":
    main()
