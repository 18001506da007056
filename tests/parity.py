"""Parity bookkeeping shared by the GPU tests, smoke() and bench.py: compares one scan's CUDA
results (labels, keys, recursion nodes) with the CPU oracle's and reports what SURVEY §8d asks
for — exact key mismatches, label agreement, normal angles over shared nodes, flips by cause."""
from __future__ import annotations

import numpy as np


def node_key(root, depth, start, n):
    return (int(root), int(depth), int(start), int(n))


def compare_nodes(gpu_nodes, orc_nodes, scan=0):
    """Matches nodes by (root, depth, start, n).  Returns a dict with counts and the normal-angle
    distribution (radians) over shared nodes that carry a plane (>= 3 inliers on both sides)."""
    g = gpu_nodes[gpu_nodes["scan"] == scan] if "scan" in gpu_nodes.dtype.names else gpu_nodes
    gk = {node_key(r["root"], r["depth"], r["start"], r["n"]): r for r in g}
    ok = {node_key(r["root"], r["depth"], r["start"], r["n"]): r for r in orc_nodes}
    shared = sorted(set(gk) & set(ok))
    angles, outcome_mismatch, inlier_diff, degenerate = [], 0, 0, 0
    worst = None
    for k in shared:
        a, b = gk[k], ok[k]
        if a["outcome"] != b["outcome"]:
            outcome_mismatch += 1
            continue
        if a["n_inliers"] >= 3 and b["n_inliers"] >= 3 and a["outcome"] in (4, 5):
            na, nb = a["normal"].astype(np.float64), b["normal"].astype(np.float64)
            # atan2(|a x b|, |a . b|): arccos of a dot product cannot resolve angles below ~3e-4 rad
            ang = float(np.arctan2(np.linalg.norm(np.cross(na, nb)), abs(float(np.dot(na, nb)))))
            angles.append(ang)
            inlier_diff += int(a["n_inliers"] != b["n_inliers"])
            if worst is None or ang > worst[0]:
                worst = (ang, k, int(a["n_inliers"]), int(b["n_inliers"]))
    angles = np.array(angles) if angles else np.zeros(0)
    return dict(n_gpu=len(gk), n_oracle=len(ok), n_shared=len(shared), only_gpu=len(set(gk) - set(ok)),
                only_oracle=len(set(ok) - set(gk)), outcome_mismatch=outcome_mismatch, inlier_count_diff=inlier_diff,
                n_planes=len(angles), max_angle=float(angles.max()) if len(angles) else 0.0,
                p999_angle=float(np.quantile(angles, 0.999)) if len(angles) else 0.0, worst=worst)


def compare_scan(labels, keys, oracle_out):
    ol, ok = oracle_out["labels"], oracle_out["keys"]
    key_mismatch = int((keys != ok).sum()) if keys is not None else -1
    agree = float((labels == ol).mean()) if len(ol) else 1.0
    diff = np.nonzero(labels != ol)[0]
    # flips by cause: class change involving the non-patch labels (2, 3) would be a binning bug
    hard = int(np.isin(labels[diff], (2, 3)).sum() + np.isin(ol[diff], (2, 3)).sum())
    return dict(n=len(ol), key_mismatch=key_mismatch, label_agreement=agree, n_flips=len(diff), n_flips_nonpatch=hard,
                ground_gpu=int((labels == 1).sum()), ground_oracle=int((ol == 1).sum()))
