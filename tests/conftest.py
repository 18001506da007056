import importlib
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

PKG_NAME = "ros2-recursive-patchwork-implementation_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def rpw():
    """The package (its directory name has hyphens, so it is imported by string)."""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the native pieces once per session."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def ref(built):
    import oracle_lib
    return oracle_lib.try_reference("strict")


@pytest.fixture(scope="session")
def gpu_handle_factory(rpw, built):
    handles = []

    def make(cfg=None, max_total_points=1 << 20, max_batch=1):
        h = rpw.Handle(cfg.to_c() if cfg is not None else None, 0, max_total_points, max_batch)
        handles.append(h)
        return h

    yield make
    for h in handles:
        h.close()
