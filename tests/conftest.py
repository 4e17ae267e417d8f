import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _build_if_missing():
    """A fresh checkout has no built artefacts (*.so is git-ignored): build the C-ABI library once, as
    __graft_entry__.build() does (nvcc cross-compiles sm_100a without a GPU).  Never a fallback: if the build is
    impossible the tests that need the library fail loudly in gfnerf_b200._lib.lib()."""
    lib = os.path.join(ROOT, "gf-nerf_b200", "libgfnerf_b200.so")
    if os.path.exists(lib):
        return
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("gf_build", os.path.join(ROOT, "gf-nerf_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    except Exception as e:  # pragma: no cover
        sys.stderr.write(f"conftest: could not build libgfnerf_b200.so ({e})\n")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    _build_if_missing()


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
