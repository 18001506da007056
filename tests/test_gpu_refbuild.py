"""The reference's OWN callers running on the B200 path.

oracle/Makefile (target `dropin`) compiles RP/test/test_recursive_patchwork.cpp and RP/src/main.cpp (+ the host
helpers they use: point_cloud_processor.cpp, lidar_fusion.cpp, visualization.cpp, cuda_interface.cu as C++) UNCHANGED
and in place from /root/reference, against an include directory in which RP/include/recursive_patchwork.hpp is
replaced by the drop-in header (host/recursive_patchwork.hpp) -- the in-place replacement INTEGRATION.md section 1
prescribes -- and links librpw_b200.so instead of RP/src/recursive_patchwork.cpp.  Eigen and OpenCV headers are absent
from the image: oracle/eigen_standin and tests/ref_build/opencv2 stand in (test infrastructure).  The binaries land in
oracle/_ref/ (git-ignored, shipped to the GPU box); /root/reference is only needed to BUILD them."""
import os
import struct
import subprocess
import zlib
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF_SRC = Path("/root/reference/src/recursive_patchwork")
REF_TEST = ROOT / "oracle" / "_ref" / "ref_test_b200"
REF_CLI = ROOT / "oracle" / "_ref" / "ref_cli_b200"


def _need(binary):
    if not binary.exists():
        if REF_SRC.exists():
            pytest.fail(f"{binary} missing although the reference sources are present: run __graft_entry__.build()")
        pytest.skip(f"{binary} not built and the reference sources are not on this machine")


def read_png(path):
    """Decoder for the 8-bit RGB, filter-0 PNGs the cv::imwrite stand-in writes: (h, w, 3) uint8 RGB."""
    data = Path(path).read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(tag + body) == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0], "PNG chunk CRC"
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 2)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    assert not raw[:, 0].any()
    return raw[:, 1:].reshape(h, w, 3)


def test_reference_callers_parse_against_the_dropin_header(built):
    """main.cpp:1-4 and test_recursive_patchwork.cpp:1-4 include recursive_patchwork.hpp, point_cloud_processor.hpp,
    lidar_fusion.hpp (Eigen::Matrix4f through the transitive <Eigen/Dense>) and visualization.hpp (cv::Mat): every one
    of the reference's caller translation units must parse with the drop-in header in place of the reference's."""
    if not REF_SRC.exists():
        pytest.skip("reference sources not on this machine")
    r = subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "syntax"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_reference_binaries_link_the_product_library(built):
    for exe in (REF_TEST, REF_CLI):
        _need(exe)
        out = subprocess.run(["readelf", "-d", str(exe)], capture_output=True, text=True).stdout
        assert "librpw_b200.so" in out
        syms = subprocess.run(["nm", "-C", str(exe)], capture_output=True, text=True).stdout
        # the reference's own segmentation translation unit is NOT in the binary: no host fitPlaneAndSplit to fall back to
        assert "fitPlaneAndSplit" not in syms and "U rpw_segment_clouds" in syms
        # ... and is not stale: every rpw_* symbol it needs is one the library in the tree exports
        need = {ln.split()[-1] for ln in subprocess.run(["nm", "-D", "--undefined-only", str(exe)], capture_output=True, text=True).stdout.splitlines()
                if ln.split() and ln.split()[-1].startswith("rpw_")}
        have = {ln.split()[-1] for ln in subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "ros2-recursive-patchwork-implementation_b200" / "librpw_b200.so")],
                                                        capture_output=True, text=True).stdout.splitlines() if ln.split()}
        assert need and need <= have, sorted(need - have)


def test_reference_test_binary_fails_loudly_without_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _need(REF_TEST)
    r = subprocess.run([str(REF_TEST)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_reference_test_suite_passes_on_the_b200_path(built):
    """RP/test/test_recursive_patchwork.cpp:166-184, unchanged, asserts live (no -DNDEBUG): testBasicFunctionality,
    testEnhancedFiltering, testPointCloudProcessor, testLidarFusion, testPerformance."""
    _need(REF_TEST)
    r = subprocess.run([str(REF_TEST)], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert "All testing passed. Good to go." in r.stdout
    for marker in ("Basic functionality", "Filtering test passed", "Pcloud processor test passed", "LiDAR fusion works!", "Performance test completed"):
        assert marker in r.stdout, marker


@pytest.mark.gpu
def test_reference_cli_runs_on_the_b200_path(built, tmp_path):
    """RP/src/main.cpp, unchanged (built without USE_ROS2: its --demo mode), through filterGroundPoints,
    sampleGroundAndObstacles and the BEV writers: exit status 0, both PNGs written, ground drawn green and
    non-ground red, counts add up."""
    _need(REF_CLI)
    r = subprocess.run([str(REF_CLI), "--demo", "--use-patchwork", "--separate-display", "--bev-width", "120", "--bev-height", "120",
                        "--x-min", "-60", "--y-min", "-60"], capture_output=True, text=True, timeout=300, cwd=tmp_path)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "Processing completed" in r.stdout
    counts = {k: int(line.split(":")[1]) for line in r.stdout.splitlines() for k in ("Ground points", "Non-ground points", "Total points")
              if line.startswith(k + ":")}
    assert counts["Total points"] == 10000
    assert counts["Ground points"] + counts["Non-ground points"] == 10000 and counts["Ground points"] > 5000 and counts["Non-ground points"] > 1000
    img = read_png(tmp_path / "demo_frame_patchwork.png")
    assert img.shape == (120, 120, 3)
    green = (img == (0, 255, 0)).all(-1).sum()
    red = (img == (255, 0, 0)).all(-1).sum()
    assert green > 1000 and red > 300 and green + red + (img == 0).all(-1).sum() == 120 * 120
    enhanced = read_png(tmp_path / "demo_frame_enhanced.png")
    assert enhanced.shape == (120, 120, 3) and enhanced.any()
