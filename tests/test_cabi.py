"""The C-ABI library loads on a machine without a GPU and exports every symbol include/gfnerf_b200.h declares
(no compute calls here).  Also: the product package never routes through the oracle."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "gfnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from gfnerf_b200 import _lib
    names = declared_functions()
    assert len(names) >= 20, names
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/gfnerf_b200.h but not exported: {missing}"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.EXPORTS) == names, sorted(set(_lib.EXPORTS) ^ set(names))


def test_version_and_error_string_without_a_gpu():
    from gfnerf_b200 import _lib
    L = _lib.lib()
    assert b"sm_100a" in L.gf_version()
    assert L.gf_mlp_param_count(64) == 11603
    assert L.gf_mlp_param_count(48) == -1
    # argument validation happens before any CUDA call
    rc = L.gf_hash_forward(-1, None, 1, 16, None, None, None, None, None, None, 0, None, None, None)
    assert rc == -1 and b"gf_hash_forward" in L.gf_last_error()


def test_library_is_sm100a_only():
    from gfnerf_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_package_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "gf-nerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libgf_oracle" not in text, f
