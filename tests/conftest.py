"""pytest configuration: registers the `gpu` marker, puts the repo root on sys.path and
aliases the (dash-named) product package as `ccb200`."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG_NAME = "chunk-compaction-in-vectorized-execution-simd_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_pkg():
    if "ccb200" in sys.modules:
        return sys.modules["ccb200"]
    pkg = importlib.import_module(PKG_NAME)
    sys.modules["ccb200"] = pkg
    return pkg


@pytest.fixture(scope="session")
def ccb():
    """The product package with the device initialised (GPU tests only)."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pkg = load_pkg()
    pkg.init(0)
    return pkg


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    oracle_lib.lib()
    return oracle_lib
