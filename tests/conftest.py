import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spmv-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "perf: prints a timing for information; never asserts on wall-clock time")


# Parity first: the GPU run is `pytest -x`, so a failure further down must never hide the SpMV parity results.
_ORDER = ["test_gpu_parity", "test_gpu_dropin", "test_gpu_fullscale", "test_gpu_group", "test_gpu_layout_build",
          "test_gpu_cg", "test_gpu_sanitizer"]


def pytest_collection_modifyitems(session, config, items):
    def key(item):
        name = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        return _ORDER.index(name) if name in _ORDER else len(_ORDER)
    items.sort(key=key)  # stable: the order inside a file is kept


@pytest.fixture(scope="session")
def oracle():
    import oracle_api
    return oracle_api.OracleLib()


@pytest.fixture(scope="session")
def spmvb():
    import spmvb as _s
    _s.build_library()
    _s.lib()
    return _s
