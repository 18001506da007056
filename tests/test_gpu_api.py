"""GPU tests of the boundary's behaviour: degenerate inputs, strides, batches, capacity errors,
the device-resident entry point, the clouds the reference returns, and the Python mirror class."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h(gpu_handle_factory):
    return gpu_handle_factory(None, 1 << 19, 8)


def test_native_library_is_the_one_running(rpw, h):
    before = h.kernel_launches()
    h.segment(rpw.synth.testsuite_cloud(1, 1000))
    assert h.kernel_launches() - before == 11  # bin, offsets, scatter, fit roots x7 size classes, fit levels


def test_degenerate_inputs(rpw, h, oracle):
    cfg = rpw.PatchworkConfig()
    h.set_config(cfg.to_c())
    assert len(h.segment(np.zeros((0, 3), np.float32))) == 0
    for pts in (np.array([[3, 4, 0], [5, 1, 0.1]], np.float32),          # < 3 in-zone points
                np.full((5, 3), np.nan, np.float32),                      # nothing survives cleaning
                np.array([[1e4, 0, 0]] * 4, np.float32),                  # all beyond the radius
                np.array([[0.5, 0.1, 0], [80.0, 0, 0], [10, -1e-9, 0], [10, 1, 0], [11, 1, 0]], np.float32),
                np.array([[10, 1, 0], [10, 1, 0], [10, 1, 0], [10, 1, 0]], np.float32)):  # identical points
        got = h.segment(pts)
        want = oracle.run(cfg, pts)["labels"]
        assert np.array_equal(got, want), (pts, got, want)


def test_stride_12_and_16_agree(rpw, h):
    pts = rpw.synth.spinning_scan(1003)
    a = h.segment(pts)
    b = h.segment(np.ascontiguousarray(pts[:, :3]))
    assert np.array_equal(a, b)


def test_batch_equals_singles_and_ragged_sizes(rpw, h):
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    h.set_config(cfg.to_c())
    scans = [rpw.synth.spinning_scan(1100, 64, 500), rpw.synth.testsuite_cloud(9, 7), np.zeros((0, 4), np.float32),
             rpw.synth.spinning_scan(1101, 64, 1875), rpw.synth.testsuite_cloud(10, 4097)]
    singles = [h.segment(s) if len(s) else np.zeros(0, np.uint8) for s in scans]
    batch = h.segment_batch(scans)
    for a, b in zip(singles, batch):
        assert np.array_equal(a, b)


def test_capacity_and_argument_errors(rpw, gpu_handle_factory):
    small = gpu_handle_factory(None, 1000, 2)
    with pytest.raises(rpw.RpwError) as e:
        small.segment(rpw.synth.testsuite_cloud(1, 2000))
    assert e.value.code == rpw.capi.RPW_ERR_CAPACITY
    with pytest.raises(rpw.RpwError) as e:
        small.segment_batch([rpw.synth.testsuite_cloud(1, 10)] * 3)
    assert e.value.code == rpw.capi.RPW_ERR_CAPACITY
    bad = rpw.PatchworkConfig(num_sectors=0)
    with pytest.raises(rpw.RpwError) as e:
        small.set_config(bad.to_c())
    assert e.value.code == rpw.capi.RPW_ERR_BAD_ARG
    # the handle still works afterwards
    assert len(small.segment(rpw.synth.testsuite_cloud(1, 900))) == 900


def test_device_resident_entry_point(rpw, h):
    import torch
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    h.set_config(cfg.to_c())
    scans = [rpw.synth.spinning_scan(1200 + i, 64, 600) for i in range(3)]
    want = h.segment_batch(scans)
    off = np.zeros(4, np.uint64)
    off[1:] = np.cumsum([len(s) for s in scans])
    d_pts = torch.from_numpy(np.concatenate(scans)).cuda()
    d_lab = torch.full((int(off[-1]),), 255, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    h.set_stream(s.cuda_stream)
    h.segment_device(d_pts.data_ptr(), off, d_lab.data_ptr())
    s.synchronize()
    h.set_stream(None)
    got = d_lab.cpu().numpy()
    assert np.array_equal(got, np.concatenate(want))


def test_clouds_follow_reference_order(rpw, h, oracle):
    cfg = rpw.PatchworkConfig(filtering_radius=40.0)
    h.set_config(cfg.to_c())
    pts = rpw.synth.testsuite_cloud(77, 6000)
    g, ng, labels = h.segment_clouds(pts)
    o = oracle.run(cfg, pts)["labels"]
    assert np.array_equal(labels, o)
    p = pts[:, :3]
    assert np.array_equal(g, p[o == 1])
    assert np.array_equal(ng, np.concatenate([p[o == 0], p[o == 2]]))


def _expected_clouds(p, labels):
    """The reference's assembly (RP/src/recursive_patchwork.cpp:402-419) from per-point labels."""
    return p[labels == 1], np.concatenate([p[labels == 0], p[labels == 2]])


def test_device_result_assembly_batch(rpw, h, oracle):
    """SURVEY section 8a row a13 on the device: stable compaction of a whole batch (ragged scans, an
    empty scan, non-finite points, beyond-radius points), host and device destinations."""
    import torch
    cfg = rpw.PatchworkConfig(filtering_radius=30.0)
    h.set_config(cfg.to_c())
    scans = [rpw.synth.testsuite_cloud(200 + i, n) for i, n in enumerate((9000, 1, 0, 4097, 12000, 4096, 33))]
    scans[0][5:9, 0] = np.nan
    scans[4][100, 2] = np.inf
    labels = h.segment_batch(scans)
    for a, l in zip(scans, labels):
        if len(a):
            assert np.array_equal(l, oracle.run(cfg, a)["labels"])
    counts = [len(a) for a in scans]
    clouds = h.last_clouds(counts)
    for a, l, (g, ng) in zip(scans, labels, clouds):
        eg, eng = _expected_clouds(a[:, :3], l)
        assert np.array_equal(g.view(np.uint32), eg.view(np.uint32)) and np.array_equal(ng.view(np.uint32), eng.view(np.uint32))
    # device path in, device clouds out
    off = np.zeros(len(scans) + 1, np.uint64); off[1:] = np.cumsum(counts)
    total = int(off[-1])
    d_pts = torch.from_numpy(np.concatenate([a for a in scans if len(a)])).cuda()
    d_lab = torch.empty(total, dtype=torch.uint8, device="cuda")
    d_g = torch.full((total, 3), -7.0, dtype=torch.float32, device="cuda")
    d_ng = torch.full((total, 3), -7.0, dtype=torch.float32, device="cuda")
    st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
    h.segment_device(d_pts.data_ptr(), off, d_lab.data_ptr())
    cnt = h.last_clouds_device(d_g.data_ptr(), d_ng.data_ptr(), len(scans))
    st.synchronize(); h.set_stream(None)
    gh, ngh = d_g.cpu().numpy(), d_ng.cpu().numpy()
    for b, (a, l) in enumerate(zip(scans, labels)):
        eg, eng = _expected_clouds(a[:, :3], l)
        o = int(off[b])
        assert tuple(cnt[b]) == (len(eg), len(eng))
        assert np.array_equal(gh[o:o + len(eg)], eg) and np.array_equal(ngh[o:o + len(eng)], eng)
        assert np.all(gh[o + len(eg):int(off[b + 1])] == -7.0)  # nothing written past a scan's cloud


def test_device_result_assembly_fused(rpw, h, oracle):
    """Clouds of a fused multi-LiDAR frame come out in vehicle coordinates, ego points removed:
    equal to the reference's fusion (oracle.fuse) followed by its assembly."""
    from test_oracle import sensor_frames
    clouds, yaws, ego = sensor_frames(rpw, seed=11)
    cfg = rpw.PatchworkConfig()
    h.set_config(cfg.to_c())
    labels = np.concatenate(h.segment_fused(clouds, yaws, ego))
    (g, ng), = h.last_clouds([sum(map(len, clouds))])
    fused, src = oracle.fuse(clouds, yaws, ego)
    eg, eng = _expected_clouds(fused[:, :3], labels[src])
    assert np.array_equal(g.view(np.uint32), eg.view(np.uint32))
    assert np.array_equal(ng.view(np.uint32), eng.view(np.uint32))


def test_sample_ground_and_obstacles_on_device(rpw, h, oracle):
    """SURVEY section 8f row 2 (RP/src/recursive_patchwork.cpp:428-465): obstacles exactly as the reference's
    loops select them from the non-ground cloud, ground context = 2000 distinct ground points (the reference
    draws them unseeded, so only membership and count are comparable)."""
    cfg = rpw.PatchworkConfig(filtering_radius=60.0)
    h.set_config(cfg.to_c())
    pts = rpw.synth.spinning_scan(1500, 64, 600)
    labels = h.segment(pts)
    o = oracle.run(cfg, pts)["labels"]
    assert (labels == o).mean() >= 0.999
    p = pts[:, :3]
    eg, eng = _expected_clouds(p, labels)
    for target, tol in ((1.1, 0.5), (0.3, 0.25)):
        sample, obstacles = h.sample_ground_and_obstacles(len(pts), target, tol, 2.5, 2000, seed=7)
        d = np.sqrt(eng[:, 0] * eng[:, 0] + eng[:, 1] * eng[:, 1], dtype=np.float32)
        want = eng[(d > np.float32(2.5)) & (np.abs(eng[:, 2] - np.float32(target)) <= np.float32(tol))]
        assert np.array_equal(obstacles, want) and len(want) > 0
        assert len(sample) == min(2000, len(eg))
        as_rows = lambda a: {r.tobytes() for r in np.ascontiguousarray(a)}
        assert len(as_rows(sample)) == len(sample) or len(as_rows(eg)) < len(eg)  # distinct draws (unless the cloud itself repeats a point)
        assert as_rows(sample) <= as_rows(eg)
    again, _ = h.sample_ground_and_obstacles(len(pts), 1.1, 0.5, 2.5, 2000, seed=7)
    first, _ = h.sample_ground_and_obstacles(len(pts), 1.1, 0.5, 2.5, 2000, seed=7)
    assert np.array_equal(again, first)  # a seed makes the draw reproducible
    # fewer ground points than the sample size: all of them, in order; no non-ground points: the whole ground cloud
    small = rpw.synth.testsuite_cloud(61, 900)
    lab = h.segment(small)
    sg, so = h.sample_ground_and_obstacles(len(small), 1.1, 0.5, 2.5, 2000, seed=3)
    assert np.array_equal(sg, small[:, :3][lab == 1])
    flat = np.zeros((500, 3), np.float32)
    flat[:, 0] = np.linspace(3.0, 3.4, 500); flat[:, 1] = np.linspace(0.1, 0.2, 500); flat[:, 2] = 0.01  # one flat patch
    lab = h.segment(flat)
    assert (lab == 1).all()
    sg, so = h.sample_ground_and_obstacles(len(flat), 1.1, 0.5, 2.5, 100, seed=3)
    assert np.array_equal(sg, flat) and len(so) == 0


def test_degenerate_cloud_order_is_the_reference(rpw, h, ref):
    """Fewer than three points inside the radius: the reference returns ({}, cleaned points) in plain input order, the
    beyond-radius points where they stood (RP/src/recursive_patchwork.cpp:339-341); K4 must not move them to the tail."""
    if ref is None:
        pytest.skip("oracle/_ref/libref_strict.so not built")
    cfg = rpw.PatchworkConfig(filtering_radius=50.0)
    h.set_config(cfg.to_c())
    rng = np.random.default_rng(5)
    for n_in in (0, 1, 2, 3):
        far = rng.uniform(60.0, 300.0, (9000, 3)).astype(np.float32) * rng.choice([-1.0, 1.0], (9000, 3)).astype(np.float32)
        pts = far.copy()
        pts[rng.choice(9000, 40, replace=False), 0] = np.nan
        where = np.sort(rng.choice(9000, n_in, replace=False))
        pts[where] = rng.uniform(2.0, 20.0, (n_in, 3)).astype(np.float32)
        r = ref.run(cfg, pts)
        g, ng, labels = h.segment_clouds(pts)
        assert np.array_equal(labels, r["labels"]), n_in
        assert np.array_equal(g.view(np.uint32), r["ground"].view(np.uint32)), n_in
        assert np.array_equal(ng.view(np.uint32), r["non_ground"].view(np.uint32)), n_in


def test_single_scan_graph_path(rpw, gpu_handle_factory, oracle):
    """Single host-path scans run as one captured CUDA graph per handle: same labels as the plain launches, one capture
    for a stream of frames of varying size, a new capture when the configuration / solver changes or a frame outgrows
    the captured grids -- and the debug switches fall back to plain launches."""
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    hd = gpu_handle_factory(cfg, 1 << 19, 2)
    frames = [rpw.synth.spinning_scan(4000 + k, 64, 800 + 13 * k) for k in range(6)]   # 51 k .. 55 k points
    want = [oracle.run(cfg, a)["labels"] for a in frames]
    hd.scan_graph(0)
    plain = [hd.segment(a) for a in frames]
    assert hd.scan_graph() == (0, 0)
    hd.scan_graph(1)
    for rep in range(3):
        for a, p_, w in zip(frames, plain, want):
            got = hd.segment(a)
            assert np.array_equal(got, p_) and (got == w).mean() >= 0.999
    launches, captures = hd.scan_graph()
    assert launches == 18 and captures == 1, (launches, captures)
    big = rpw.synth.spinning_scan(4100, 64, 1875)                                         # 120 k points: outgrows the graph
    assert (hd.segment(big) == oracle.run(cfg, big)["labels"]).mean() >= 0.999
    assert hd.scan_graph() == (19, 2)
    assert np.array_equal(hd.segment(frames[0]), plain[0]) and hd.scan_graph() == (20, 2)  # smaller frames reuse the larger graph
    cfg2 = rpw.PatchworkConfig(filtering_radius=80.0, num_sectors=12)
    hd.set_config(cfg2.to_c())
    assert (hd.segment(frames[1]) == oracle.run(cfg2, frames[1])["labels"]).mean() >= 0.999
    assert hd.scan_graph() == (21, 3)
    hd.set_plane_solver(rpw.capi.SOLVER_REFERENCE)
    assert np.array_equal(hd.segment(frames[2]), oracle.run(cfg2, frames[2])["labels"]) and hd.scan_graph() == (22, 4)
    hd.set_plane_solver(rpw.capi.SOLVER_HYBRID)
    hd.enable_nodes(True)                                                                 # node records: plain launches
    hd.segment(frames[3])
    assert len(hd.debug_nodes()) > 0 and hd.scan_graph()[0] == 22
    hd.enable_nodes(False)
    two = hd.segment_batch([frames[0], frames[1]])                                        # batches: plain launches
    assert hd.scan_graph()[0] == 22 and (two[1] == oracle.run(cfg2, frames[1])["labels"]).mean() >= 0.999
    # clouds, sample and raster of the last single scan work after a graph launch
    g, ng, lab = hd.segment_clouds(frames[4])
    assert hd.scan_graph()[0] == 23 and len(g) == int((lab == 1).sum()) and len(ng) == int(np.isin(lab, (0, 2)).sum())
    hd.close()


def test_reserve_grows_the_handle_in_place(rpw, gpu_handle_factory, oracle):
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    hd = gpu_handle_factory(cfg, 1 << 14, 1)
    hd.set_plane_solver(rpw.capi.SOLVER_EIGEN_QR)
    small = rpw.synth.testsuite_cloud(9, 9000)
    big = rpw.synth.spinning_scan(1234, 64, 900)
    want_small, want_big = oracle.run(cfg, small)["labels"], oracle.run(cfg, big)["labels"]
    assert np.array_equal(hd.segment(small), want_small)
    with pytest.raises(rpw.RpwError) as e:
        hd.segment(big)
    assert e.value.code == rpw.capi.RPW_ERR_CAPACITY
    hd.reserve(len(big) + 2 * len(small), 4)
    assert hd.capacity() == (len(big) + 2 * len(small), 4)
    hd.reserve(100, 1)  # never shrinks
    assert hd.capacity() == (len(big) + 2 * len(small), 4)
    assert (hd.segment(big) == want_big).mean() >= 0.9999
    got = hd.segment_batch([small, big, small])
    assert np.array_equal(got[0], want_small) and np.array_equal(got[2], want_small)
    hd.close()


def test_sample_ground_and_obstacles_is_the_reference(rpw, h, ref):
    """f2 pinned to the reference itself: RecursivePatchwork::sampleGroundAndObstacles as compiled from
    RP/src/recursive_patchwork.cpp:428-465 (oracle/_ref/libref_strict.so) against rpw_sample_ground_and_obstacles on
    the same scans.  The obstacle list is compared bit for bit, in order; the ground context sample (unseeded
    std::mt19937 in the reference) by count, distinctness and membership in the reference's own ground cloud."""
    if ref is None:
        pytest.skip("oracle/_ref/libref_strict.so not built")
    as_rows = lambda a: {r.tobytes() for r in np.ascontiguousarray(a)}
    cases = [(rpw.PatchworkConfig(filtering_radius=60.0), rpw.synth.spinning_scan(1500, 64, 600), (1.1, 0.5)),
             (rpw.PatchworkConfig(filtering_radius=60.0), rpw.synth.spinning_scan(1501, 64, 600), (0.3, 0.25)),
             (rpw.PatchworkConfig(), rpw.synth.testsuite_cloud(42, 10000), (1.1, 0.5)),   # the CLI's demo cloud shape, its defaults
             (rpw.PatchworkConfig(), rpw.synth.testsuite_cloud(61, 900), (1.1, 0.5)),     # fewer ground points than the sample size
             (rpw.PatchworkConfig(filtering_radius=150.0), rpw.synth.solidstate_merged(2500, 120, 100), (1.1, 0.5))]
    for cfg, pts, (target, tol) in cases:
        h.set_config(cfg.to_c())
        r = ref.run(cfg, pts)
        want = ref.sample_ground_and_obstacles(cfg, pts, target, tol)
        labels = h.segment(pts)
        assert np.array_equal(labels, r["labels"])
        sample, obstacles = h.sample_ground_and_obstacles(len(pts), target, tol, 2.5, 2000, seed=0)
        k = min(2000, len(r["ground"]))
        assert len(want) >= k
        want_sample, want_obstacles = want[:k], want[k:]
        assert np.array_equal(obstacles.view(np.uint32), want_obstacles.view(np.uint32)), (len(obstacles), len(want_obstacles))
        assert len(sample) == len(want_sample)
        ground_rows = as_rows(r["ground"])
        assert as_rows(sample) <= ground_rows and as_rows(want_sample) <= ground_rows
        if len(ground_rows) == len(r["ground"]):
            assert len(as_rows(sample)) == len(sample)
        if len(r["ground"]) <= 2000:  # no draw: the whole ground cloud in order, in both
            assert np.array_equal(sample, want_sample)


def test_bev_rasters_are_the_reference(rpw, h, ref):
    """f4 pinned to the reference itself: Visualization::createGroundNonGroundImage / createBEVImage as compiled from
    RP/src/visualization.cpp:18-80 (against the cv::Mat stand-in of tests/ref_build) fed with the reference's own
    clouds, against rpw_bev_image, pixel for pixel -- the CLI's three rasters (RP/src/main.cpp:268-300), at the CLI's
    default geometry and two others."""
    if ref is None:
        pytest.skip("oracle/_ref/libref_strict.so not built")
    cases = [(rpw.PatchworkConfig(), rpw.synth.testsuite_cloud(42, 10000)),
             (rpw.PatchworkConfig(filtering_radius=60.0), rpw.synth.spinning_scan(1600, 32, 500)),
             (rpw.PatchworkConfig(filtering_radius=80.0), rpw.synth.spinning_scan(1601, 64, 900))]
    views = [(300, 150, -150.0, -75.0, 150.0, 75.0),   # main.cpp defaults: x_min + bev_width, y_min + bev_height
             (200, 160, -40.0, -32.0, 40.0, 32.0), (64, 48, -10.0, 0.0, 30.0, 30.0)]
    painted = 0
    for cfg, pts in cases:
        h.set_config(cfg.to_c())
        r = ref.run(cfg, pts)
        labels = h.segment(pts)
        assert np.array_equal(labels, r["labels"])
        g, ng = r["ground"], r["non_ground"]
        for (w, hh, x0, y0, x1, y1) in views:
            assert np.array_equal(h.bev_image(h.BEV_CLASSES, w, hh, x0, y0, x1, y1), ref.bev(0, g, ng, w, hh, x0, y0, x1, y1))
            assert np.array_equal(h.bev_image(h.BEV_HEIGHT_NONGROUND, w, hh, x0, y0, x1, y1), ref.bev(1, ng, None, w, hh, x0, y0, x1, y1))
            got = h.bev_image(h.BEV_HEIGHT_ALL, w, hh, x0, y0, x1, y1)
            assert np.array_equal(got, ref.bev(1, np.concatenate([g, ng]), None, w, hh, x0, y0, x1, y1))
            painted += int(got.any())
    assert painted == len(cases) * len(views)


def _bev_reference(clouds_colours, width, height, x_min, y_min, x_max, y_max):
    """Sequential drawing of RP/src/visualization.cpp (createBEVImage / createGroundNonGroundImage): float scale
    factors, int truncation, later points overwrite earlier ones."""
    img = np.zeros((height, width, 3), np.uint8)
    xs = np.float32(width) / (np.float32(x_max) - np.float32(x_min))
    ys = np.float32(height) / (np.float32(y_max) - np.float32(y_min))
    for pts, colour in clouds_colours:
        x = ((pts[:, 0] - np.float32(x_min)) * xs).astype(np.int32)   # C++ static_cast<int>: truncation
        y = ((pts[:, 1] - np.float32(y_min)) * ys).astype(np.int32)
        ok = (x >= 0) & (x < width) & (y >= 0) & (y < height)
        col = colour(pts) if callable(colour) else np.broadcast_to(np.array(colour, np.uint8), (len(pts), 3))
        for xi, yi, c in zip(x[ok], y[ok], col[ok]):  # in order: the last point on a pixel wins
            img[yi, xi] = c
    return img


def test_bev_rasters_on_device(rpw, h, oracle):
    """SURVEY section 8f row 4 (RP/src/visualization.cpp:18-113): the three rasters the CLI writes, drawn on the
    device from the device-resident clouds, pixel for pixel equal to the reference's sequential drawing."""
    cfg = rpw.PatchworkConfig(filtering_radius=60.0)
    h.set_config(cfg.to_c())
    pts = rpw.synth.spinning_scan(1600, 32, 500)
    labels = h.segment(pts)
    eg, eng = _expected_clouds(pts[:, :3], labels)

    def height_colour(p):
        i = np.minimum(np.float32(255.0), np.maximum(np.float32(0.0), (p[:, 2] + np.float32(2.0)) * np.float32(50.0))).astype(np.int32)
        return np.stack([i, i, np.full_like(i, 255)], 1).astype(np.uint8)

    for (w, hh, x0, y0, x1, y1) in ((200, 160, -40.0, -32.0, 40.0, 32.0), (64, 48, -10.0, 0.0, 30.0, 30.0)):
        got = h.bev_image(h.BEV_CLASSES, w, hh, x0, y0, x1, y1)
        assert np.array_equal(got, _bev_reference([(eg, (0, 255, 0)), (eng, (0, 0, 255))], w, hh, x0, y0, x1, y1))
        got = h.bev_image(h.BEV_HEIGHT_NONGROUND, w, hh, x0, y0, x1, y1)
        assert np.array_equal(got, _bev_reference([(eng, height_colour)], w, hh, x0, y0, x1, y1))
        got = h.bev_image(h.BEV_HEIGHT_ALL, w, hh, x0, y0, x1, y1)
        assert np.array_equal(got, _bev_reference([(eg, height_colour), (eng, height_colour)], w, hh, x0, y0, x1, y1))
        assert got.any()
    with pytest.raises(rpw.RpwError):
        h.bev_image(7, 10, 10, 0, 0, 1, 1)


def test_changing_workload_on_one_handle(rpw, gpu_handle_factory, oracle):
    """The fit kernels' grids are sized from the previous call's per-class patch counts; results must not
    depend on how wrong that estimate is.  One handle sees very different calls in a row (a batch of
    spinning scans, one big-patch frame, a tiny cloud, a deep-recursion scan, an empty call, the batch
    again); every call is compared with the oracle."""
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    hd = gpu_handle_factory(cfg, 1 << 20, 8)
    batch = [rpw.synth.spinning_scan(7000 + i, 32, 700) for i in range(6)]
    big = rpw.synth.solidstate_merged(2500, 300, 200)
    tiny = rpw.synth.testsuite_cloud(71, 40)
    deep = rpw.synth.spinning_scan(3300, 64, 1024, 1)
    want = {id(a): oracle.run(cfg, a)["labels"] for a in batch + [big, tiny, deep]}

    def check(clouds):
        got = hd.segment_batch(clouds)
        for a, l in zip(clouds, got):
            assert (l == want[id(a)]).mean() >= 0.999, len(a)

    for _ in range(2):
        check(batch)
        check([big])
        check([tiny])
        check([deep])
        assert len(hd.segment(np.zeros((0, 3), np.float32))) == 0
        check([tiny, big, batch[0], deep])
    hd.close()


def test_python_mirror_class(rpw, oracle):
    cfg = rpw.PatchworkConfig(sensor_height=1.2, filtering_radius=50.0, num_sectors=8, max_iter=50)  # testBasicFunctionality's config
    rp = rpw.RecursivePatchwork(cfg, max_points=1 << 16)
    pts = rpw.synth.testsuite_cloud(5, 5000)[:, :3]
    ground, non_ground = rp.filterGroundPoints(pts)
    # the reference test's own assertions (RP/test/test_recursive_patchwork.cpp:74-76)
    assert len(ground) + len(non_ground) <= len(pts) and len(ground) > 0 and len(non_ground) > 0
    o = oracle.run(cfg, pts)["labels"]
    assert len(ground) == int((o == 1).sum()) and len(non_ground) == int(np.isin(o, (0, 2)).sum())
    g0, n0 = rp.filterGroundPoints(np.zeros((0, 3), np.float32))
    assert len(g0) == 0 and len(n0) == 0
    rp.setConfig(rpw.PatchworkConfig())
    assert rp.getConfig().filtering_radius == 150.0
    assert len(rp.cleanPoints(np.array([[1, 2, 3], [np.nan, 0, 0]], np.float32))) == 1
    rp.close()


def test_set_config_changes_sector_count(rpw, h, oracle):
    pts = rpw.synth.testsuite_cloud(12, 8000)
    for S in (10, 24, 3, 10):
        cfg = rpw.PatchworkConfig(num_sectors=S, filtering_radius=70.0)
        h.set_config(cfg.to_c())
        got = h.segment(pts)
        keys = h.debug_keys(len(pts))
        o = oracle.run(cfg, pts)
        assert np.array_equal(keys, o["keys"])
        assert (got == o["labels"]).mean() >= 0.999


def test_pointcloud2_buffer_ingest(rpw, h, oracle):
    """x, y, z read in place from a PointCloud2-style record buffer
    (point_step 32, fields at 4 / 12 / 20 like an x-y-z-intensity-ring layout with padding)."""
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    h.set_config(cfg.to_c())
    pts = rpw.synth.spinning_scan(1300, 64, 400)
    n = len(pts)
    rec = np.zeros((n, 8), np.float32)
    rec[:, 1] = pts[:, 0]; rec[:, 3] = pts[:, 1]; rec[:, 5] = pts[:, 2]
    rec[:, 0] = 123.0; rec[:, 2] = np.nan; rec[:, 7] = -1.0   # junk in the other fields must not matter
    got = h.segment_pc2(rec.tobytes(), n, 32, 4, 12, 20)
    want = oracle.run(cfg, pts)["labels"]
    assert np.array_equal(got, want)
    # wide records with xyz in front through the plain entry point (the bag loader's convention)
    wide = np.zeros((n, 6), np.float32)
    wide[:, :3] = pts[:, :3]
    assert np.array_equal(h.segment(wide), want)
    with pytest.raises(rpw.RpwError):
        h.segment_pc2(rec.tobytes(), n, 30, 4, 12, 20)
    with pytest.raises(rpw.RpwError):
        h.segment_pc2(rec.tobytes(), n, 32, 4, 12, 30)


def test_fused_multi_lidar_frame(rpw, h, oracle):
    """rotation + ego removal + concatenation folded into the binning kernel.
    Expected labels: oracle fusion -> oracle segmentation of the merged cloud, mapped back to the
    sensors' own points; ego-removed points carry label 4."""
    from test_oracle import sensor_frames
    clouds, yaws, ego = sensor_frames(rpw, seed=9)
    cfg = rpw.PatchworkConfig()
    h.set_config(cfg.to_c())
    labels, st = h.segment_fused(clouds, yaws, ego, want_stats=True)
    fused, src = oracle.fuse(clouds, yaws, ego)
    want = np.full(sum(map(len, clouds)), 4, np.uint8)
    want[src] = oracle.run(cfg, fused)["labels"]
    got = np.concatenate(labels)
    assert np.array_equal(got == 4, want == 4), "ego removal differs"
    assert np.array_equal(np.isin(got, (2, 3)), np.isin(want, (2, 3)))
    assert (got == want).mean() >= 0.999
    assert st.n_points == len(want) and st.n_ground == int((got == 1).sum())
    # one sensor, no rotation, zero ego radius == the plain entry point
    plain = h.segment(clouds[0])
    one = h.segment_fused(clouds[:1], [0.0], [-1.0])[0]
    assert np.array_equal(plain, one)


def test_concurrent_handles_with_recursion(rpw, oracle):
    """Several handles on one device, each from its own thread and stream, on a scene that recurses
    (so every handle's level kernel really runs its grid barrier): no deadlock, same labels."""
    import threading
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    pts = rpw.synth.spinning_scan(3000, 128, 1024, 1)
    want = oracle.run(cfg, pts)["labels"]
    assert oracle.run(cfg, pts, want_nodes=True)["stats"]["n_splits"] > 5
    results, errors = {}, []

    def work(k):
        try:
            hk = rpw.Handle(cfg.to_c(), 0, len(pts) + 16, 1)
            for _ in range(15):
                results[k] = hk.segment(pts)
            hk.close()
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
        assert not t.is_alive(), "a handle is stuck"
    assert not errors
    for k in range(4):
        assert (results[k] == want).mean() >= 0.999
