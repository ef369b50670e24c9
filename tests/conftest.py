"""pytest configuration: `gpu` marker, import paths, shared helpers."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nvalchemi-toolkit-ops_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The CUDA library is built in-tree and git-ignored: (re)build it when it is missing or older than its sources
    (nvcc cross-compiles sm_100a without a GPU, about a minute), so the suite does not depend on a previous build()."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("nvnl_build_script", os.path.join(ROOT, "nvalchemi-toolkit-ops_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.needs_build():
        try:
            mod.build()
        except Exception as e:  # pragma: no cover - reported by the tests that need the library
            print(f"[conftest] could not build the CUDA library: {e}", file=sys.stderr)


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
