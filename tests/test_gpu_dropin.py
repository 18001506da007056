"""Builds tests/cpp/test_dropin.cpp against the C++ drop-in header + librpw_b200.so and runs it on
the GPU: the reference's own smoke tests (RP/test/test_recursive_patchwork.cpp:51-98,146-164)
re-stated, plus an exact clouds-vs-oracle check through the C++ entry point."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "ros2-recursive-patchwork-implementation_b200"


def build_dropin(tmp_path):
    exe = tmp_path / "test_dropin"
    cxx = shutil.which("g++") or "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", f"-I{ROOT / 'include'}", f"-I{PKG / 'host'}", str(ROOT / "tests" / "cpp" / "test_dropin.cpp"),
                    "-o", str(exe), f"-L{PKG}", "-lrpw_b200", f"-Wl,-rpath,{PKG}"], check=True)
    return exe


def test_dropin_header_compiles_against_the_abi(built, tmp_path):
    """CPU-side: the reference-compatible header + C-ABI link cleanly (no run)."""
    assert build_dropin(tmp_path).exists()


@pytest.mark.gpu
def test_reference_style_cpp_tests_pass_on_gpu(rpw, built, oracle, tmp_path):
    exe = build_dropin(tmp_path)
    cfg = rpw.PatchworkConfig(filtering_radius=80.0)
    pts = np.ascontiguousarray(rpw.synth.spinning_scan(1005)[:, :3])
    labels = oracle.run(cfg, pts)["labels"]
    pts.tofile(tmp_path / "cloud.bin")
    labels.tofile(tmp_path / "labels.bin")
    r = subprocess.run([str(exe), str(tmp_path / "cloud.bin"), str(tmp_path / "labels.bin"), "80"], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "ALL DROP-IN TESTS PASSED" in r.stdout
