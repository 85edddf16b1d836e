import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "host_emul")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build native code once per session (no-op when up to date)."""
    import subprocess
    from dwarf_p_cloudsc_b200 import build
    build.build_all()
    targets = ["emul", "unit"]
    if os.path.exists(os.path.join(ROOT, "oracle", "cloudsc_oracle.c")):
        targets.append("port")
    if os.path.exists("/root/reference/src/cloudsc_c/cloudsc/cloudsc_c.c"):
        targets += ["ref", "ref32", "reffast", "cpuref", "level0"]
    if os.path.exists("/root/reference/src/cloudsc_cuda/cloudsc/cloudsc_c_k_caching.cu"):
        targets.append("gpuref")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), *targets], check=True, stdout=subprocess.DEVNULL)
    return True


@pytest.fixture(scope="session")
def oracle(built):
    import pyoracle
    return pyoracle.best_oracle()


@pytest.fixture(scope="session")
def input_data():
    from dwarf_p_cloudsc_b200 import params, state
    cols, scalars, klon, klev = state.load_input()
    return cols, scalars, klon, klev
