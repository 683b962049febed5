import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _ensure_built():
    # the .so files are git-ignored build artefacts; build them once per checkout (nvcc cross-compiles without a GPU)
    lib = os.path.join(ROOT, "dopt_b200", "lib", "libdopt_b200.so")
    host = os.path.join(ROOT, "dopt_b200", "lib", "libdopt_host.so")
    if not os.path.exists(lib) or (os.path.isdir(os.path.join(ROOT, "dopt_b200", "host")) and not os.path.exists(host)):
        import __graft_entry__
        __graft_entry__.build()
    yield
