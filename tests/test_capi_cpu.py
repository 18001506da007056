"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/rpw_b200.h declares, mirrors the reference's config, and fails loudly without a GPU
(no compute calls are made here)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    text = (ROOT / "include" / "rpw_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rpw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(rpw, built):
    lib = rpw.load_library()
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"librpw_b200.so does not export {name}"
    assert sorted(rpw.capi.EXPORTS) == declared
    assert lib.rpw_abi_version() == 1


def test_default_config_mirrors_reference_struct(rpw, built):
    c = rpw.capi.default_config()
    d = rpw.PatchworkConfig()
    # RP/include/recursive_patchwork.hpp:25-36
    assert (c.sensor_height, c.max_range, c.num_sectors, c.max_iter, c.adaptive_seed_height) == (pytest.approx(1.2), 150.0, 10, 100, 1)
    assert (c.th_seeds, c.th_dist, c.th_outlier, c.filtering_radius, c.max_split_depth) == (
        pytest.approx(0.15), pytest.approx(0.2), pytest.approx(0.08), 150.0, 1000)
    assert rpw.PatchworkConfig.from_c(d.to_c()) == rpw.PatchworkConfig.from_c(c)
    assert C.sizeof(rpw.capi.RpwConfig) == 40


def test_node_record_layout(rpw):
    assert rpw.capi.NODE_DTYPE.itemsize == 72


def test_no_cpu_fallback(rpw, built):
    """Without a CUDA device the product must refuse to work, not quietly compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rpw.RpwError) as e:
        rpw.Handle(None, 0, 1 << 16, 1)
    assert e.value.code == rpw.capi.RPW_ERR_NO_DEVICE
    with pytest.raises(rpw.RpwError):
        rpw.RecursivePatchwork()


def test_bad_arguments_are_rejected_before_any_device_work(rpw, built):
    lib = rpw.load_library()
    h = C.c_void_p()
    cfg = rpw.capi.default_config()
    cfg.num_sectors = 0
    assert lib.rpw_create(C.byref(cfg), 0, 1000, 1, C.byref(h)) == rpw.capi.RPW_ERR_BAD_ARG
    cfg.num_sectors = 10
    assert lib.rpw_create(C.byref(cfg), 0, 0, 1, C.byref(h)) == rpw.capi.RPW_ERR_BAD_ARG
    assert lib.rpw_create(C.byref(cfg), 0, 1000, 1, None) == rpw.capi.RPW_ERR_BAD_ARG
    assert lib.rpw_segment(None, None, 0, 12, None, None) == rpw.capi.RPW_ERR_BAD_ARG


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = ROOT / "ros2-recursive-patchwork-implementation_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.hpp")) + list(pkg.rglob("*.cpp")):
        t = p.read_text()
        assert "oracle_lib" not in t and "librpw_oracle" not in t and "rpwo_" not in t and "libref_" not in t, p
