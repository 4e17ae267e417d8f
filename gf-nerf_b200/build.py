"""Builds the C-ABI library (gf-nerf_b200/libgfnerf_b200.so) with nvcc for sm_100a, in-tree.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the
resulting .so travels to the GPU box with the repo snapshot.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgfnerf_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fopenmp",      # host loops of the octree builder (csrc/octree_build.cu)
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "gfnerf_b200.h")]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        # static cudart (nvcc default): independent of whichever libcudart torch brings along
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fopenmp",
                               "-o", LIB] + objs)
    return LIB


TORCH_LIB = os.path.join(HERE, "f2nerf_bindings_b200.so")


def build_torch_bindings(force: bool = False) -> str:
    """The TorchScript custom classes (torch.classes.my_classes.*) over the C-ABI: csrc/torch_bindings.cpp compiled with
    g++ against libtorch, linked to libgfnerf_b200.so through an $ORIGIN rpath.  Loaded with
    torch.classes.load_library(TORCH_LIB), exactly how the reference loads its f2nerf-bindings.so."""
    import torch
    from torch.utils import cpp_extension
    src = os.path.join(CSRC, "torch_bindings.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "gfnerf_b200.h"), LIB]
    if not (force or _stale(TORCH_LIB, deps)):
        return TORCH_LIB
    inc = []
    for p in cpp_extension.include_paths(device_type="cuda") if "device_type" in cpp_extension.include_paths.__code__.co_varnames \
            else cpp_extension.include_paths(cuda=True):
        inc += ["-I", p]
    lib_dir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cxx11 = int(torch.compiled_with_cxx11_abi())
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={cxx11}", "-I", "/usr/local/cuda/include"] + inc + [
        src, "-o", TORCH_LIB, "-L", HERE, "-lgfnerf_b200", "-L", lib_dir, "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10",
        "-lc10_cuda", f"-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{lib_dir}"]
    subprocess.check_call(cmd)
    return TORCH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--torch" in sys.argv:
        print(build_torch_bindings(force="--force" in sys.argv))
