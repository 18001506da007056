"""Parity soak, gated (round 1 ran it as a script): many seeded scans of every shape through the C-ABI against the CPU
oracle.  Bars: ring/sector keys bit-exact on every scan; in the reference-order mode (RPW_SOLVER_REFERENCE) labels are
IDENTICAL on every scan; with the default solver every scan is >= 99.9 % except where a long-running fit amplifies the
summation order (the two known stress seeds are listed, and asserted to be fixed by replaying long fits)."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu

BAR = 0.999
# seeds whose default-solver agreement is below the bar because of one bistable / never-converging patch
# (profiles/parity_soak_r01.md): they must pass with rpw_set_exact_replay(8) and be identical in the reference-order mode
KNOWN_STRESS = {("C5", 3116), ("C4", 2121)}


def shapes(rpw):
    return [
        ("C2", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.spinning_scan(s), range(5000, 5064)),
        ("C2nan", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.spinning_scan(s, nan_per_million=5000), range(6000, 6008)),
        ("C5", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.dense_urban_scan(s), range(3100, 3148)),
        ("C4", rpw.PatchworkConfig(), lambda s: rpw.synth.solidstate_merged(s), range(2100, 2124)),
        ("C1", rpw.PatchworkConfig(), lambda s: rpw.synth.testsuite_cloud(s, 10000), range(100, 132)),
        ("C1pct", rpw.PatchworkConfig(adaptive_seed_height=False, num_sectors=37), lambda s: rpw.synth.testsuite_cloud(s, 10000), range(300, 316)),
    ]


@pytest.mark.parametrize("shape", ["C2", "C2nan", "C5", "C4", "C1", "C1pct"])
def test_soak(shape, rpw, built):
    name, cfg, gen, seeds = next(x for x in shapes(rpw) if x[0] == shape)
    oracle = oracle_lib.Oracle()
    ocfg = oracle_lib.to_cfg(cfg)
    with ThreadPoolExecutor(16) as ex:
        scans = list(ex.map(gen, seeds))
        want = list(ex.map(lambda a: oracle.run(ocfg, a), scans))
    h = rpw.Handle(cfg.to_c(), 0, max(len(a) for a in scans) + 4096, 1)
    rows = []
    try:
        for seed, a, o in zip(seeds, scans, want):
            h.set_plane_solver(rpw.capi.SOLVER_HYBRID); h.set_exact_replay(-1)
            fast = h.segment(a)
            keys_bad = int((h.debug_keys(len(a)) != o["keys"]).sum())
            h.set_exact_replay(8)
            replay = h.segment(a)
            h.set_exact_replay(-1); h.set_plane_solver(rpw.capi.SOLVER_REFERENCE)
            exact = h.segment(a)
            rows.append((seed, len(a), keys_bad, int((fast != o["labels"]).sum()), int((replay != o["labels"]).sum()), int((exact != o["labels"]).sum())))
    finally:
        h.close()
    n = sum(r[1] for r in rows)
    print(f"{name}: scans {len(rows)} points {n} key mismatches {sum(r[2] for r in rows)} labels differing: default {sum(r[3] for r in rows)} "
          f"replay(8) {sum(r[4] for r in rows)} reference-order {sum(r[5] for r in rows)}; worst scan default "
          f"{min(1 - r[3] / r[1] for r in rows):.6f} replay(8) {min(1 - r[4] / r[1] for r in rows):.6f}")
    for seed, npts, kb, df, dr, de in rows:
        assert kb == 0, (name, seed, "keys")
        assert de == 0, (name, seed, f"{de} labels differ in the reference-order mode")
        assert dr <= (1 - BAR) * npts, (name, seed, f"replay(8): {dr} of {npts}")
        if (name, seed) not in KNOWN_STRESS:
            assert df <= (1 - BAR) * npts, (name, seed, f"default solver: {df} of {npts}")
