"""Where a scan's default-solver labels leave the reference's: recursion nodes of the default run against those of the
reference-order run (same kernel family, bit-identical to the oracle), matched by (root patch, depth, start, size).

    python tools/gpu_node_diff.py C5 31245 [replay_K]
"""
import importlib, sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    shape, seed = sys.argv[1], int(sys.argv[2])
    replay = int(sys.argv[3]) if len(sys.argv) > 3 else -1
    rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
    gen = {"C2": rpw.synth.spinning_scan, "C4": rpw.synth.solidstate_merged, "C5": rpw.synth.dense_urban_scan}[shape]
    cfg = rpw.synth.config_for(shape)
    a = gen(seed)
    h = rpw.Handle(cfg.to_c(), 0, len(a) + 4096, 1)
    h.enable_nodes(True)
    h.set_plane_solver(rpw.capi.SOLVER_REFERENCE)
    want = h.segment(a); ref_nodes = h.debug_nodes()
    h.set_plane_solver(rpw.capi.SOLVER_HYBRID); h.set_exact_replay(replay)
    got = h.segment(a); nodes = h.debug_nodes()
    print(f"{shape} seed {seed}: {len(a)} points, labels differing {int((got != want).sum())} ({(got == want).mean():.5f} agree), "
          f"nodes {len(nodes)} vs {len(ref_nodes)} in the reference-order run")
    key = lambda r: (int(r["root"]), int(r["depth"]), int(r["start"]), int(r["n"]))
    R = {key(r): r for r in ref_nodes}
    G = {key(r): r for r in nodes}
    only_g = sorted(set(G) - set(R)); only_r = sorted(set(R) - set(G))
    print(f"nodes only in the default run {len(only_g)}, only in the reference-order run {len(only_r)}")
    rows = []
    for k in sorted(set(G) & set(R), key=lambda k: (k[1], k[0])):
        g, r = G[k], R[k]
        if int(g["outcome"]) != int(r["outcome"]) or int(g["n_inliers"]) != int(r["n_inliers"]):
            rows.append((k, g, r))
    print(f"shared nodes with a different outcome or inlier count: {len(rows)} (shallowest first)")
    for k, g, r in rows[:25]:
        print(f"  root {k[0]} depth {k[1]} n {k[3]}: outcome {int(g['outcome'])}/{int(r['outcome'])} iters {int(g['iters'])}/{int(r['iters'])} "
              f"inliers {int(g['n_inliers'])}/{int(r['n_inliers'])} residual {float(g['residual']):.6g}/{float(r['residual']):.6g} "
              f"normal-angle {np.degrees(np.arccos(np.clip(np.dot(g['normal'], r['normal']), -1, 1))):.4f} deg")
    h.close()


if __name__ == "__main__":
    main()
