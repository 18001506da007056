"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded
inputs and against the golden vectors made from the reference itself.

Bars (BASELINE.json north star): zone / ring / sector keys bit-exact; per-point ground labels
agree on >= 99.9 % of points; plane normals within 1e-3 rad on every node both sides share."""
import numpy as np
import pytest

import golden_util
import parity

pytestmark = pytest.mark.gpu

LABEL_BAR = 0.999
NORMAL_BAR = 1e-3


@pytest.fixture(scope="module")
def h(gpu_handle_factory):
    hd = gpu_handle_factory(None, 1 << 20, 8)
    hd.enable_nodes(True)
    return hd


def run_case(h, oracle, cfg, pts):
    h.set_config(cfg.to_c())
    labels = h.segment(pts)
    keys = h.debug_keys(len(pts))
    nodes = h.debug_nodes()
    o = oracle.run(cfg, pts, want_nodes=True)
    return labels, keys, nodes, o


@pytest.mark.parametrize("name", golden_util.names())
def test_golden_vectors(name, rpw, h, oracle):
    g = golden_util.load(name, rpw.PatchworkConfig)
    h.set_config(g["cfg"].to_c())
    labels = h.segment(g["points"])
    agree = float((labels == g["labels"]).mean())
    assert agree >= LABEL_BAR, f"{name}: agreement {agree}"
    # class of non-patch points is exact
    assert np.array_equal(np.isin(labels, (2, 3)), np.isin(g["labels"], (2, 3)))
    assert np.array_equal(labels[np.isin(labels, (2, 3))], g["labels"][np.isin(g["labels"], (2, 3))])


CASES = {
    "C1_3000_default": lambda rpw: (rpw.PatchworkConfig(), rpw.synth.testsuite_cloud(42, 3000)),
    "C1_5000_r50_s8": lambda rpw: (rpw.PatchworkConfig(filtering_radius=50.0, num_sectors=8, max_iter=50), rpw.synth.testsuite_cloud(43, 5000)),
    "C1_10000_default": lambda rpw: (rpw.PatchworkConfig(), rpw.synth.testsuite_cloud(44, 10000)),
    "C1_10000_splits": lambda rpw: (rpw.PatchworkConfig(), rpw.synth.testsuite_cloud(42, 10000)),
    "C1_nonadaptive": lambda rpw: (rpw.PatchworkConfig(adaptive_seed_height=False), rpw.synth.testsuite_cloud(42, 10000)),
    "C1_37_sectors": lambda rpw: (rpw.PatchworkConfig(num_sectors=37, filtering_radius=60.0), rpw.synth.testsuite_cloud(47, 20000)),
    "C1_max_iter_3": lambda rpw: (rpw.PatchworkConfig(max_iter=3), rpw.synth.testsuite_cloud(48, 10000)),
    "C1_max_iter_0": lambda rpw: (rpw.PatchworkConfig(max_iter=0), rpw.synth.testsuite_cloud(48, 10000)),
    "C1_depth_limit_1": lambda rpw: (rpw.PatchworkConfig(max_split_depth=1), rpw.synth.testsuite_cloud(42, 10000)),
    "C2_120k": lambda rpw: (rpw.PatchworkConfig(filtering_radius=80.0), rpw.synth.spinning_scan(1000)),
    "C2_120k_other_seed": lambda rpw: (rpw.PatchworkConfig(filtering_radius=80.0), rpw.synth.spinning_scan(1777, nan_per_million=5000)),
    "C4_300k": lambda rpw: (rpw.PatchworkConfig(), rpw.synth.solidstate_merged(2000)),
    "C5_262k_deep": lambda rpw: (rpw.PatchworkConfig(filtering_radius=80.0), rpw.synth.dense_urban_scan(3000)),
    "C5_262k_deep_b": lambda rpw: (rpw.PatchworkConfig(filtering_radius=80.0), rpw.synth.dense_urban_scan(3001)),
}


@pytest.mark.parametrize("solver", ["hybrid", "eigen_qr"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_parity_against_oracle(case, solver, rpw, h, oracle):
    """Every scene with the default solver (hybrid) and with the reference's own QR sequence."""
    cfg, pts = CASES[case](rpw)
    h.set_plane_solver(rpw.capi.SOLVER_HYBRID if solver == "hybrid" else rpw.capi.SOLVER_EIGEN_QR)
    try:
        labels, keys, nodes, o = run_case(h, oracle, cfg, pts)
    finally:
        h.set_plane_solver(rpw.capi.SOLVER_HYBRID)
    rep = parity.compare_scan(labels, keys, o)
    nrep = parity.compare_nodes(nodes, o["nodes"])
    print(case, rep, nrep)
    assert rep["key_mismatch"] == 0, "ring/sector/zone keys must be bit-exact"
    assert rep["n_flips_nonpatch"] == 0
    assert rep["label_agreement"] >= LABEL_BAR
    # every node the two implementations share: same outcome, normals within the bar
    assert nrep["n_shared"] >= 0.98 * nrep["n_oracle"]
    assert nrep["outcome_mismatch"] <= max(1, nrep["n_shared"] // 100)
    assert nrep["max_angle"] <= NORMAL_BAR, nrep


def _nodes_bit_identical(gpu_nodes, orc_nodes):
    """Every oracle node exists on the GPU with the same outcome, iteration count, inlier count, split axis, the same BITS of
    centroid, normal and residual, and the same split value."""
    gk = {parity.node_key(r["root"], r["depth"], r["start"], r["n"]): r for r in gpu_nodes}
    ok = {parity.node_key(r["root"], r["depth"], r["start"], r["n"]): r for r in orc_nodes}
    assert set(gk) == set(ok), (len(gk), len(ok))
    bad = []
    for k, b in ok.items():
        a = gk[k]
        same = all(int(a[f]) == int(b[f]) for f in ("outcome", "iters", "n_inliers", "split_axis"))
        if same and b["outcome"] in (4, 5):
            for f in ("centroid", "normal", "residual"):
                same = same and np.array_equal(np.asarray(a[f], np.float32).view(np.uint32), np.asarray(b[f], np.float32).view(np.uint32))
            # the split value by VALUE: among coordinates that compare equal (-0.0 and +0.0) a sort may put either at the
            # median position (std::sort / qsort leave it unspecified, the device's radix select orders by bits); the
            # partition `v <= median` cannot tell them apart
            same = same and float(a["median"]) == float(b["median"])
        if not same:
            bad.append((k, {f: (a[f], b[f]) for f in ("outcome", "iters", "n_inliers", "residual")}))
    assert not bad, bad[:5]
    return len(ok)


@pytest.mark.parametrize("case", sorted(CASES))
def test_reference_order_mode_is_bit_identical(case, rpw, h, oracle):
    """RPW_SOLVER_REFERENCE: sequential float sums in the reference's order + Eigen's QR sequence.  Not a tolerance
    test: labels, keys and every recursion node's outcome / iterations / inliers / centroid / normal / residual bits equal
    the oracle's (which is pinned bit for bit to the reference's own strict-IEEE build)."""
    cfg, pts = CASES[case](rpw)
    h.set_plane_solver(rpw.capi.SOLVER_REFERENCE)
    try:
        labels, keys, nodes, o = run_case(h, oracle, cfg, pts)
    finally:
        h.set_plane_solver(rpw.capi.SOLVER_HYBRID)
    assert np.array_equal(keys, o["keys"])
    n_diff = int((labels != o["labels"]).sum())
    assert n_diff == 0, f"{n_diff} labels differ"
    _nodes_bit_identical(nodes, o["nodes"])


@pytest.mark.parametrize("name", golden_util.names())
def test_golden_vectors_reference_order_mode(name, rpw, h):
    """The fixtures generated from the reference itself (tests/golden/make_golden.py): identical labels."""
    g = golden_util.load(name, rpw.PatchworkConfig)
    h.set_config(g["cfg"].to_c())
    h.set_plane_solver(rpw.capi.SOLVER_REFERENCE)
    try:
        labels = h.segment(g["points"])
    finally:
        h.set_plane_solver(rpw.capi.SOLVER_HYBRID)
    assert np.array_equal(labels, g["labels"])


def test_exact_replay_of_long_fits(rpw, h, oracle):
    """Selective form (rpw_set_exact_replay(8) on the default solver): only fits of more than 8 iterations are redone in
    the reference's order.  On the two stress scans of the round-1 soak that fell below the bar (a bistable inner-ring patch
    in C5 seed 3116, a never-converging 15 k-point patch in C4 seed 2121) it restores >= 99.9 %."""
    for cfg, pts in ((rpw.PatchworkConfig(filtering_radius=80.0), rpw.synth.dense_urban_scan(3116)),
                     (rpw.PatchworkConfig(), rpw.synth.solidstate_merged(2121))):
        h.set_config(cfg.to_c())
        want = oracle.run(cfg, pts)["labels"]
        fast = float((h.segment(pts) == want).mean())
        h.set_exact_replay(8)
        try:
            replayed = float((h.segment(pts) == want).mean())
        finally:
            h.set_exact_replay(-1)
        print(f"n={len(pts)} fast {fast:.6f} replay(8) {replayed:.6f}")
        assert replayed >= LABEL_BAR and replayed >= fast


@pytest.mark.parametrize("case", ["C2_120k", "C4_300k", "C5_262k_deep", "C5_262k_deep_b", "C1_nonadaptive", "C1_10000_splits"])
def test_cluster_fit_agrees_with_single_block_fit(case, rpw, built, oracle, monkeypatch):
    """Calls of one or two scans spread patches above 4096 points over thread-block clusters (distributed shared memory,
    rpw_fit_cluster_kernel); batches keep a patch on one block.  Same algorithm, the moments summed in another tree order:
    both forms must meet the bar against the oracle, make the same node decisions (outcome, split axis, median) and agree with
    each other to the same bar."""
    cfg, pts = CASES[case](rpw)
    if case.startswith("C1"):  # small clouds: widen the patches so that some exceed 4096 points
        cfg.num_sectors = 1
    o = oracle.run(cfg, pts, want_nodes=True)
    got = {}
    for profile in ("0", "1"):
        monkeypatch.setenv("RPW_FIT_PROFILE", profile)
        hd = rpw.Handle(cfg.to_c(), 0, len(pts) + 4096, 1)
        hd.enable_nodes(True)
        labels = hd.segment(pts)
        nodes = hd.debug_nodes()
        hd.enable_nodes(False)
        again = hd.segment(pts)  # the captured graph
        assert np.array_equal(again, labels)
        hd.close()
        rep = parity.compare_scan(labels, None, o)
        nrep = parity.compare_nodes(nodes, o["nodes"])
        print(case, "profile", profile, rep, nrep)
        assert rep["label_agreement"] >= LABEL_BAR and nrep["max_angle"] <= NORMAL_BAR
        assert nrep["outcome_mismatch"] <= max(1, nrep["n_shared"] // 100)
        got[profile] = (labels, nodes)
    big = got["1"][1][got["1"][1]["n"] > 4096]
    assert len(big) > 0, "no patch large enough for the cluster kernels in this case"
    assert (got["0"][0] == got["1"][0]).mean() >= LABEL_BAR


def test_cluster_fit_seed_fallback_and_ties(rpw, built, oracle, monkeypatch):
    """The rare seed path on a cluster node: a big patch (> 4096 points) with no point below z_th takes its three LOWEST points
    as seeds (RP/src/recursive_patchwork.cpp:173-181); with ties in z across the cut libstdc++'s heap-select is replayed by one
    thread, which then reads the other blocks' points through distributed shared memory.  Labels must equal the oracle's in both
    forms (these slabs are clean planes: no borderline points)."""
    rng = np.random.default_rng(11)
    n = 9000
    ang = rng.uniform(0.05, 0.55, n)            # one sector of ten
    rad = rng.uniform(18.0, 26.0, n)            # one ring
    for ties in (False, True):
        z = 4.0 + 0.02 * rad + rng.normal(0, 0.004, n)   # a tilted slab well above sensor_height + 0.2 * rel
        if ties:
            z = np.round(z, 2)                  # many equal heights, also among the lowest
            z[rng.choice(n, 7, replace=False)] = z.min()
        pts = np.stack([rad * np.cos(ang), rad * np.sin(ang), z], 1).astype(np.float32)
        cfg = rpw.PatchworkConfig(filtering_radius=80.0)
        o = oracle.run(cfg, pts, want_nodes=True)
        assert o["nodes"]["n"].max() > 4096
        for profile in ("0", "1"):
            monkeypatch.setenv("RPW_FIT_PROFILE", profile)
            hd = rpw.Handle(cfg.to_c(), 0, n + 4096, 1)
            labels = hd.segment(pts)
            hd.close()
            assert np.array_equal(labels, o["labels"]), (ties, profile, int((labels != o["labels"]).sum()))


def test_deep_recursion_is_exercised(rpw, h, oracle):
    """C5 must actually recurse (demonstrated, not assumed)."""
    cfg, pts = CASES["C5_262k_deep_b"](rpw)
    labels, keys, nodes, o = run_case(h, oracle, cfg, pts)
    assert o["stats"]["max_depth"] >= 4 and o["stats"]["n_splits"] >= 15
    assert nodes["depth"].max() == o["stats"]["max_depth"]
    assert int((nodes["outcome"] == 5).sum()) == o["stats"]["n_splits"]


def test_device_atan2_is_host_libm(h, oracle):
    import ctypes as C
    rng = np.random.default_rng(3)
    y = rng.uniform(-150, 150, 400000).astype(np.float32)
    x = rng.uniform(-150, 150, 400000).astype(np.float32)
    y[:1000] = 0.0
    x[1000:2000] = 0.0
    x[2000:3000] = 1.0
    y[3000:4000] *= 1e-20
    dev = h.debug_atan2(y, x)
    host = oracle.atan2_restated(y[:60000], x[:60000])  # restated == libm is proven on CPU
    assert np.array_equal(dev[:60000].view(np.uint32), host.view(np.uint32))
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.max(np.abs(dev.astype(np.float64) - ref)) < 1e-6


def test_device_eigensolver_bit_exact(h, oracle):
    rng = np.random.default_rng(4)
    A = rng.normal(size=(20000, 3, 3)).astype(np.float32)
    A = A @ A.transpose(0, 2, 1)
    A[:10000, 2, :] *= 1e-2
    A[:10000, :, 2] *= 1e-2
    A = ((A + A.transpose(0, 2, 1)) * 0.5).astype(np.float32)
    A[0] = 0
    A[1] = np.eye(3)
    ev, vec = h.debug_eig3(A)
    oev, ovec = oracle.eig3(A)
    assert np.array_equal(ev.view(np.uint32), oev.view(np.uint32))
    assert np.array_equal(vec.view(np.uint32), ovec.view(np.uint32))


def test_full_size_properties(rpw, h):
    """Size-independent properties at BASELINE's full sizes (no oracle in the loop):
    labels are a partition; permuting the input permutes non-patch classes and keeps counts of
    beyond/dropped; translating z of an all-ground flat cloud keeps it all ground."""
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    h.set_config(cfg.to_c())
    pts = rpw.synth.dense_urban_scan(3003)
    labels, st = h.segment(pts, want_stats=True)
    assert st.n_points == len(pts) and st.n_ground + st.n_nonground + st.n_beyond + st.n_dropped == len(pts)
    fin = np.isfinite(pts[:, :3]).all(axis=1)
    assert np.array_equal(labels == 3, ~fin)
    d = np.sqrt((pts[:, 0] * pts[:, 0] + pts[:, 1] * pts[:, 1]).astype(np.float32))
    assert np.array_equal(labels[fin] == 2, d[fin] > np.float32(80.0))
    # idempotence: same input, same labels
    assert np.array_equal(h.segment(pts), labels)
    # a perfectly flat disc: every patch takes the z-range early-out -> all binned points ground
    rng = np.random.default_rng(5)
    r = rng.uniform(1.5, 70, 200000).astype(np.float32)
    a = rng.uniform(0, 2 * np.pi, 200000).astype(np.float32)
    flat = np.stack([r * np.cos(a), r * np.sin(a), np.full_like(r, 0.3)], 1).astype(np.float32)
    lf = h.segment(flat)
    keys = h.debug_keys(len(flat))
    assert np.all(lf[keys < 0xFFFD] == 1)


def test_restructured_qr_is_the_generic_qr(h):
    """The latency-restructured solver the fit kernel runs must return the bits of the generic
    Eigen-sequence solver (which test_device_eigensolver_bit_exact pins to the oracle)."""
    rng = np.random.default_rng(11)
    n = 20000
    A = rng.normal(size=(n, 3, 3)).astype(np.float32)
    A = A @ A.transpose(0, 2, 1)
    A[: n // 2, 2, :] *= 1e-2
    A[: n // 2, :, 2] *= 1e-2
    A[0] = 0
    A[1] = np.eye(3)
    A[2] = np.diag([1.0, 1.0, 0.0])
    A[3, :, :] = np.outer([1, 2, 3], [1, 2, 3])  # rank 1
    sc = np.stack([A[:, 0, 0], A[:, 1, 0], A[:, 1, 1], A[:, 2, 0], A[:, 2, 1], A[:, 2, 2]], 1).astype(np.float32)
    # plane-like scatter matrices (what the fit feeds the solver), tiny and huge scales, exact zeros
    planes = []
    for i in range(20000):
        ext = rng.uniform(0.5, 30, 2)
        p = rng.normal(size=(int(rng.integers(3, 400)), 3)) * np.array([ext[0], ext[1], rng.uniform(0.0, 0.2)])
        a, b = rng.uniform(-0.3, 0.3, 2)
        p[:, 2] += a * p[:, 0] + b * p[:, 1]
        d = (p - p.mean(0)).astype(np.float32)
        S = (d.T @ d) / np.float32(len(p) - 1)
        planes.append([S[0, 0], S[1, 0], S[1, 1], S[2, 0], S[2, 1], S[2, 2]])
    planes = np.array(planes, np.float32)
    planes[:200] *= np.float32(1e-30)
    planes[200:400] *= np.float32(1e30)
    planes[400:600, 3:5] = 0
    planes[600:700, 1] = 0
    planes[700:800] *= np.float32(1e-38)
    sc = np.concatenate([sc, planes])
    generic, _ = h.debug_normal(sc, 1)
    fast, _ = h.debug_normal(sc, 2)
    assert np.array_equal(generic.view(np.uint32), fast.view(np.uint32))
    spec, _ = h.debug_normal(sc, 3)  # branch-free arithmetic + IEEE fallback: what the fit kernel runs
    assert np.array_equal(generic.view(np.uint32), spec.view(np.uint32))
    raw, _ = h.debug_normal(sc, 4)   # without the fallback: in-range solves must already be exact, and most are in range
    inr = raw[:, 2] != 2.0
    assert np.array_equal(generic[inr].view(np.uint32), raw[inr].view(np.uint32))
    assert inr[-len(planes) + 800:].mean() > 0.99


def test_closed_form_solver_accuracy(h):
    rng = np.random.default_rng(12)
    mats = []
    for i in range(3000):
        n = int(rng.integers(3, 3000))
        ext = rng.uniform(0.5, 30, 2)
        p = rng.normal(size=(n, 3)) * np.array([ext[0], ext[1], rng.uniform(0.0, 0.2)])
        a, b = rng.uniform(-0.3, 0.3, 2)
        p[:, 2] += a * p[:, 0] + b * p[:, 1]
        d = (p - p.mean(0)).astype(np.float32)
        S = d.T @ d
        mats.append([S[0, 0], S[1, 0], S[1, 1], S[2, 0], S[2, 1], S[2, 2]])
    # two smallest eigenvalues close to each other (Newton on the cubic meets a nearly double root): a
    # seed set seen on the C5 stress scene (eigenvalues 1419.6, 1453.0, 4160.9) and rotated diag(1, 1+g, 3)
    mats.append([3264.94, 1280.3462, 2328.6538, -30.985556, -41.74468, 1439.9283])
    for g in (3e-1, 1e-1, 3e-2, 1e-2, 3e-3, 1e-3, 3e-4):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        S = q @ np.diag([1.0, 1.0 + g, 3.0]) @ q.T
        mats.append([S[0, 0], S[1, 0], S[1, 1], S[2, 0], S[2, 1], S[2, 2]])
    mats = np.array(mats, np.float32)
    full = np.zeros((len(mats), 3, 3))
    full[:, 0, 0] = mats[:, 0]; full[:, 1, 0] = full[:, 0, 1] = mats[:, 1]; full[:, 1, 1] = mats[:, 2]
    full[:, 2, 0] = full[:, 0, 2] = mats[:, 3]; full[:, 2, 1] = full[:, 1, 2] = mats[:, 4]; full[:, 2, 2] = mats[:, 5]
    w, v = np.linalg.eigh(full)
    ref = v[:, :, 0]
    nrm, _ = h.debug_normal(mats, 0)
    nd = nrm.astype(np.float64)
    ang = np.arctan2(np.linalg.norm(np.cross(nd, ref), axis=1), np.abs((nd * ref).sum(1)))
    ok = (w[:, 1] - w[:, 0]) / np.abs(w).max(1) > 1e-4
    assert ang[ok].max() < 1e-6
    assert np.all(nrm[:, 2] >= 0) and np.allclose(np.linalg.norm(nrm, axis=1), 1, atol=1e-6)


@pytest.mark.parametrize("case", ["C1_10000_splits", "C2_120k", "C4_300k", "C5_262k_deep"])
def test_closed_form_solver_label_parity(case, rpw, gpu_handle_factory, oracle):
    """The faster solver is not the reference's rounding, so the tree of fits may differ in chaotic
    patches (a plane fit sitting between two layers tips over on a 1e-6 rad change).  All scenes,
    the two-layer stress scene C5 included, clear the 99.9 % bar; the default solver reproduces them
    exactly (test_parity_against_oracle)."""
    hd = gpu_handle_factory(None, 1 << 19, 1)
    hd.set_plane_solver(rpw.capi.SOLVER_CLOSED_FORM)
    cfg, pts = CASES[case](rpw)
    hd.set_config(cfg.to_c())
    labels = hd.segment(pts)
    keys = hd.debug_keys(len(pts))
    o = oracle.run(cfg, pts)
    rep = parity.compare_scan(labels, keys, o)
    print(case, rep)
    assert rep["key_mismatch"] == 0 and rep["n_flips_nonpatch"] == 0
    assert rep["label_agreement"] >= LABEL_BAR


def test_sector_edges_take_the_exact_path(rpw, h, oracle):
    """K1 decides sectors from a cheap angle estimate unless the point is within 5e-5 rad of a sector
    edge; those fall back to the bit-exact libm sequence.  Put points ON the edges (and a few ulps
    around them, all four quadrants, +-0 coordinates) and demand exact keys."""
    import oracle_lib
    for S, R in ((10, 80.0), (37, 60.0), (128, 150.0), (1, 50.0)):
        cfg = rpw.PatchworkConfig(num_sectors=S, filtering_radius=R)
        _, delta = oracle.zone_model(oracle_lib.to_cfg(cfg))
        edges = (np.arange(S + 1, dtype=np.float32) * np.float32(delta)).astype(np.float32)
        ang = []
        for k in range(-6, 7):
            a = edges.copy()
            for _ in range(abs(k)):
                a = np.nextafter(a, np.float32(np.inf if k > 0 else -np.inf)).astype(np.float32)
            ang.append(a)
        ang = np.concatenate(ang + [edges + np.float32(4e-5), edges - np.float32(4e-5), edges + np.float32(6e-5)])
        rng = np.random.default_rng(S)
        r = rng.uniform(1.5, R * 0.99, len(ang)).astype(np.float32)
        pts = np.stack([r * np.cos(ang.astype(np.float64)), r * np.sin(ang.astype(np.float64)), np.zeros_like(r)], 1).astype(np.float32)
        extra = np.array([[5, 0.0, 0], [5, -0.0, 0], [-5, 0.0, 0], [-5, -0.0, 0], [0.0, 5, 0], [-0.0, 5, 0], [0.0, -5, 0], [5, -1e-9, 0],
                          [5, 1e-9, 0], [5, -1e-30, 0], [1.0, 0, 0], [R, 0, 0]], np.float32)
        pts = np.concatenate([pts, extra])
        h.set_config(cfg.to_c())
        h.segment(pts)
        keys = h.debug_keys(len(pts))
        o = oracle.run(cfg, pts)
        assert np.array_equal(keys, o["keys"]), (S, np.nonzero(keys != o["keys"])[0][:10])


def _one_patch_cloud(n, seed, two_layers):
    """Everything inside one ring/sector patch of the default zone model (ring 6, sector 0)."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(45.0, 75.0, n)
    a = rng.uniform(0.02, 0.60, n)
    z = rng.normal(0.0, 0.03, n)
    if two_layers:
        z = z + (rng.uniform(size=n) < 0.5) * 0.95
    return np.stack([r * np.cos(a), r * np.sin(a), z], 1).astype(np.float32)


def _dup_xy(rpw):
    p = rpw.synth.testsuite_cloud(54, 40000)[:, :3].copy()
    p[:, :2] = np.round(p[:, :2] * np.float32(2)) / np.float32(2)  # x, y on a 0.5 m grid; z stays continuous
    return p


def _quantised_z(rpw):
    p = rpw.synth.testsuite_cloud(57, 30000)[:, :3].copy()
    p[:, 2] = np.round(p[:, 2] * np.float32(8)) / np.float32(8)
    return p


def _degenerate_geometry(rpw):
    """A cloud whose patches hold collinear runs, exact duplicates and coplanar sheets: zero and rank-one covariances."""
    rng = np.random.default_rng(61)
    t = np.linspace(2.0, 55.0, 6000, dtype=np.float32)
    line = np.stack([t, np.float32(0.3) * t, np.full_like(t, 0.05)], 1)                         # one straight line
    dup = np.tile(np.array([[12.0, -7.0, 0.1]], np.float32), (3000, 1))                         # 3000 identical points
    sheet = np.stack([rng.uniform(-40, 40, 20000), rng.uniform(-40, 40, 20000), np.zeros(20000)], 1).astype(np.float32)  # z == 0 exactly
    wall = np.stack([np.full(4000, 20.0), rng.uniform(-10, 10, 4000), rng.uniform(0, 3, 4000)], 1).astype(np.float32)    # x == 20 exactly
    return np.concatenate([line, dup, sheet, wall])


def _extreme_magnitudes(rpw):
    p = rpw.synth.testsuite_cloud(62, 20000)[:, :3].copy()
    p[::97] *= np.float32(1e20)      # x*x overflows to inf: beyond the radius
    p[5::101] *= np.float32(1e-30)   # denormal-scale coordinates: inside 1 m, in no ring
    p[7::103, 2] = np.float32(3e38)  # finite but enormous height inside the zone
    p[11::107, 0] = np.float32(-0.0)
    return p


EDGE_CASES = {
    # the fixed-point loop of :185-217 cut short: one iteration, and none at all (the final fit then sees the seed mask)
    "max_iter_one": lambda rpw: (rpw.PatchworkConfig(filtering_radius=80.0, max_iter=1), rpw.synth.spinning_scan(1003, 64, 900)),
    "max_iter_zero": lambda rpw: (rpw.PatchworkConfig(filtering_radius=80.0, max_iter=0), rpw.synth.spinning_scan(1004, 64, 900)),
    "degenerate_geometry": lambda rpw: (rpw.PatchworkConfig(filtering_radius=60.0), _degenerate_geometry(rpw)),
    "extreme_magnitudes": lambda rpw: (rpw.PatchworkConfig(filtering_radius=60.0), _extreme_magnitudes(rpw)),
    # R = inf: the ring table is 1, inf, inf, ...: every point from 1 m out lands in ring 0
    "radius_infinite": lambda rpw: (rpw.PatchworkConfig(filtering_radius=float("inf")), rpw.synth.testsuite_cloud(63, 20000)),
    "percentile_seeds_many_sectors": lambda rpw: (rpw.PatchworkConfig(adaptive_seed_height=False, num_sectors=90, filtering_radius=70.0, th_seeds=0.05),
                                                  rpw.synth.spinning_scan(1005, 32, 1200)),
    # a 300k-point patch: streamed from L2 at depth 0, collapses and splits many levels deep, so the
    # radix select, the stable partition and the level kernel all run in streaming mode first
    "one_huge_patch_two_layers": lambda rpw: (rpw.PatchworkConfig(), _one_patch_cloud(300000, 1, True)),
    "one_huge_patch_flat": lambda rpw: (rpw.PatchworkConfig(), _one_patch_cloud(200000, 2, False)),
    # ring edges collapse when R <= 1: every in-zone point is unbinned (reference loop finds no ring)
    "radius_below_one": lambda rpw: (rpw.PatchworkConfig(filtering_radius=0.9), rpw.synth.testsuite_cloud(51, 4000) * np.float32(0.02)),
    "one_sector": lambda rpw: (rpw.PatchworkConfig(num_sectors=1, filtering_radius=60.0), rpw.synth.testsuite_cloud(52, 30000)),
    "128_sectors": lambda rpw: (rpw.PatchworkConfig(num_sectors=128, filtering_radius=60.0), rpw.synth.testsuite_cloud(53, 60000)),
    "no_splits_allowed": lambda rpw: (rpw.PatchworkConfig(filtering_radius=80.0, max_split_depth=0), rpw.synth.spinning_scan(3000, 128, 1024, 1)),
    # a tiny distance threshold makes almost every fit collapse: recursion limited only by
    # n >= 50 + 10 * depth and the 25 m^2 rule
    "tiny_th_dist": lambda rpw: (rpw.PatchworkConfig(filtering_radius=80.0, th_dist=0.004), rpw.synth.spinning_scan(1000, 64, 900)),
    # heavy coordinate duplication: medians with many ties, empty right children
    "duplicated_coordinates": lambda rpw: (rpw.PatchworkConfig(filtering_radius=60.0, th_dist=0.01), _dup_xy(rpw)),
    "million_points": lambda rpw: (rpw.PatchworkConfig(), rpw.synth.testsuite_cloud(55, 1000000)),
    # no seed below the threshold anywhere AND z on a coarse grid: the lowest-3 fallback meets exact ties,
    # which the device resolves by replaying libstdc++'s heap-select (std::partial_sort, :175-176)
    "lowest3_with_ties": lambda rpw: (rpw.PatchworkConfig(sensor_height=-5.0), _quantised_z(rpw)),
    "sensor_height_negative": lambda rpw: (rpw.PatchworkConfig(sensor_height=-5.0), rpw.synth.testsuite_cloud(56, 20000)),  # < 3 seeds everywhere: lowest-3 fallback
}


@pytest.mark.parametrize("case", sorted(EDGE_CASES))
def test_edge_cases_against_oracle(case, rpw, gpu_handle_factory, oracle):
    cfg, pts = EDGE_CASES[case](rpw)
    pts = np.ascontiguousarray(pts[:, :3], np.float32)
    hd = gpu_handle_factory(cfg, len(pts) + 1024, 1)
    hd.enable_nodes(True)
    labels, st = hd.segment(pts, want_stats=True)
    keys = hd.debug_keys(len(pts))
    nodes = hd.debug_nodes()
    o = oracle.run(cfg, pts, want_nodes=True)
    rep = parity.compare_scan(labels, keys, o)
    nrep = parity.compare_nodes(nodes, o["nodes"])
    print(case, rep, {k: nrep[k] for k in ("n_gpu", "n_oracle", "n_shared", "outcome_mismatch", "max_angle")}, "levels", st.n_levels)
    assert rep["key_mismatch"] == 0 and rep["n_flips_nonpatch"] == 0
    assert rep["label_agreement"] >= LABEL_BAR
    assert nrep["n_shared"] >= 0.98 * nrep["n_oracle"]
    hd.close()


def _fuzz_case(rpw, k):
    """Seeded random configuration + cloud: every live PatchworkConfig field, all three cloud generators,
    scaled / shifted so that sensor height, radius and thresholds meet the data in different regimes."""
    rng = np.random.default_rng(9000 + k)
    cfg = rpw.PatchworkConfig(
        sensor_height=float(rng.choice([1.2, 0.4, 2.0, -0.5, 0.05])),
        num_sectors=int(rng.choice([1, 2, 3, 7, 10, 16, 31, 64, 128])),
        max_iter=int(rng.choice([0, 1, 2, 5, 20, 100])),
        adaptive_seed_height=bool(rng.integers(0, 2)),
        th_seeds=float(rng.choice([0.05, 0.15, 0.5])),
        th_dist=float(rng.choice([0.02, 0.1, 0.2, 0.6])),
        filtering_radius=float(rng.choice([1.5, 12.0, 40.0, 80.0, 150.0, 400.0])),
        max_split_depth=int(rng.choice([0, 1, 3, 1000])))
    kind = int(rng.integers(0, 3))
    if kind == 0:
        pts = rpw.synth.testsuite_cloud(int(rng.integers(1, 1 << 30)), int(rng.integers(200, 40000)))
    elif kind == 1:
        pts = rpw.synth.spinning_scan(int(rng.integers(1, 1 << 30)), int(rng.choice([16, 32, 64])), int(rng.integers(100, 900)),
                                      int(rng.integers(0, 2)), int(rng.choice([0, 200, 20000])))
    else:
        pts = rpw.synth.solidstate_merged(int(rng.integers(1, 1 << 30)), int(rng.integers(40, 200)), int(rng.integers(30, 120)))
    pts = np.ascontiguousarray(pts[:, :3], np.float32).copy()
    pts *= np.float32(rng.choice([1.0, 1.0, 0.25, 3.0]))
    pts[:, 2] += np.float32(rng.choice([0.0, 0.0, 0.3, -0.3]))
    return cfg, pts


@pytest.mark.parametrize("k", range(32))
def test_randomised_configs_against_oracle(k, rpw, h, oracle):
    cfg, pts = _fuzz_case(rpw, k)
    labels, keys, nodes, o = run_case(h, oracle, cfg, pts)
    rep = parity.compare_scan(labels, keys, o)
    nrep = parity.compare_nodes(nodes, o["nodes"])
    print(k, cfg, len(pts), rep, {q: nrep[q] for q in ("n_gpu", "n_oracle", "n_shared", "outcome_mismatch", "max_angle")})
    assert rep["key_mismatch"] == 0 and rep["n_flips_nonpatch"] == 0
    assert rep["label_agreement"] >= LABEL_BAR
    assert nrep["n_shared"] >= 0.98 * nrep["n_oracle"]


@pytest.mark.parametrize("case", sorted(EDGE_CASES))
def test_edge_cases_reference_order_mode(case, rpw, gpu_handle_factory, oracle):
    """The same edge cases in the reference-order mode: labels and node records identical to the oracle's, not merely within
    the bar -- ties in the seed fallback, collapsing fits, duplicated coordinates, 128 sectors, a million points."""
    cfg, pts = EDGE_CASES[case](rpw)
    pts = np.ascontiguousarray(pts[:, :3], np.float32)
    hd = gpu_handle_factory(cfg, len(pts) + 1024, 1)
    hd.set_plane_solver(rpw.capi.SOLVER_REFERENCE)
    hd.enable_nodes(True)
    labels = hd.segment(pts)
    nodes = hd.debug_nodes()
    o = oracle.run(cfg, pts, want_nodes=True)
    assert int((labels != o["labels"]).sum()) == 0
    _nodes_bit_identical(nodes, o["nodes"])
    hd.close()


@pytest.mark.parametrize("k", range(32))
def test_randomised_configs_reference_order_mode(k, rpw, h, oracle):
    cfg, pts = _fuzz_case(rpw, k)
    h.set_plane_solver(rpw.capi.SOLVER_REFERENCE)
    try:
        labels, keys, nodes, o = run_case(h, oracle, cfg, pts)
    finally:
        h.set_plane_solver(rpw.capi.SOLVER_HYBRID)
    assert np.array_equal(keys, o["keys"])
    assert int((labels != o["labels"]).sum()) == 0, (k, cfg)
    _nodes_bit_identical(nodes, o["nodes"])
