#!/usr/bin/env python
"""Pulls the DEVICE code of the reference's two kernel files out of /root/reference, where it lies, into
oracle/_ref/ref_kernels.inc (git-ignored build output) so that oracle/ref_driver.cpp can compile it for the host.

TEST INFRASTRUCTURE ONLY.  Nothing of the reference is copied into the repository: this script runs at build time
(`make -C oracle ref`, __graft_entry__.build()) in the container that has /root/reference, writes only under
oracle/_ref/, and the generated include is deleted again once the shared object is linked.

What is extracted, verbatim and in file order:
  * every top-level `__global__` / `__device__` function definition (with its `template<...>` line) of
      gfnerf/bindings/field/Hash3DAnchored_cuda.cu      Hash3DAnchored{Forward,Backward}Kernel
      gfnerf/bindings/PtsSampler/PersSampler_cuda.cu    GetIntersection, FindRayOctreeIntersectionKernel,
                                                        QueryFrameTransform{,Jac}, RayMarchKernel, GetEdgeSamplesKernel,
                                                        GetPointsAnchorsKernel..., MarkVistNodeKernel, MarkInvalidNodes,
                                                        CheckVisible, MarkInvisibleNodesKernel
  * the `#define` constants those bodies use (the two .cu files, PersSampler.h, Hash3DAnchored.h, Utils/Common.h)
  * `struct alignas(32) TransInfo / TreeNode / EdgePool` of PtsSampler/PersSampler.h.
  * two HOST member functions of PtsSampler/PersSampler.cpp, into a second include (ref_host_fns.inc):
      PersOctree::ProcOctree (:154-417, compaction / path compression / subdivision of the node array) and
      PersOctree::ConstructEdgePool (:833-895) -- plain C++ over std::vector<TreeNode> apart from a torch
      prologue / epilogue, which compiles against the stand-in Tensor of ref_shim/torch_stub.h.
The other host functions of those files (torch tensor code, `<<< >>>` launches) are NOT extracted: the launch
sequence of PersSampler::GetSamples (PersSampler_cuda.cu:321-477) is restated in ref_driver.cpp.

Why the reference's own build cannot be used: it needs un-vendored External/tiny-cuda-nn and a
locally patched External/eigen-3.4.0.  The device code itself needs neither tcnn nor torch, only a handful of
fixed-size Eigen types, which oracle/ref_shim/eigen_subset.h provides.
"""
import os
import re
import sys

REF = os.environ.get("GF_REFERENCE", "/root/reference")
BIND = os.path.join(REF, "gfnerf", "bindings")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")

DEFINE_OK = re.compile(r"^#define\s+(Wec\w+|Watrix\w+|N_PROS|N_CHANNELS|N_LEVELS|RES_\w+|PersMatType|TransWetType|"
                       r"MAX_\w+|OCC_\w+|ABS_\w+|REL_\w+|INIT_NODE_STAT|DivUp|LINEAR_IDX)\b")


def strip_comments_keep_lines(src):
    """Blank out // and /* */ comments (keeping newlines) so that brace matching is not fooled."""
    out, i, n = [], 0, len(src)
    while i < n:
        if src.startswith("//", i):
            j = src.find("\n", i)
            j = n if j < 0 else j
            out.append(" " * (j - i))
            i = j
        elif src.startswith("/*", i):
            j = src.find("*/", i)
            j = n if j < 0 else j + 2
            out.append("".join(c if c == "\n" else " " for c in src[i:j]))
            i = j
        else:
            out.append(src[i])
            i += 1
    return "".join(out)


def device_functions(path):
    """[(first_line, last_line, text)] of top-level definitions whose signature carries __global__ / __device__."""
    src = open(path, encoding="utf-8", errors="replace").read()
    clean = strip_comments_keep_lines(src)
    found, pos = [], 0
    for m in re.finditer(r"__(global|device)__", clean):
        if m.start() < pos:
            continue
        # only at brace depth 0
        if clean.count("{", 0, m.start()) != clean.count("}", 0, m.start()):
            continue
        # start of the declaration: back to the previous ';', '}' or preprocessor line end
        start = max(clean.rfind(";", 0, m.start()), clean.rfind("}", 0, m.start()))
        start = 0 if start < 0 else start + 1
        head = clean[start:m.start()]
        hash_pos = head.rfind("#")
        if hash_pos >= 0:
            start += head.find("\n", hash_pos) + 1
        open_brace = clean.find("{", m.end())
        semi = clean.find(";", m.end())
        if open_brace < 0 or (0 <= semi < open_brace):
            continue                                             # a declaration only
        depth, j = 0, open_brace
        while True:
            if clean[j] == "{":
                depth += 1
            elif clean[j] == "}":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        text = src[start:j + 1].strip("\n")
        first = src.count("\n", 0, start + (len(src[start:]) - len(src[start:].lstrip()))) + 1
        found.append((first, src.count("\n", 0, j) + 1, text))
        pos = j + 1
    return found


def host_function(path, signature_start):
    """(first_line, last_line, text) of the top-level definition that starts with `signature_start`."""
    src = open(path, encoding="utf-8", errors="replace").read()
    clean = strip_comments_keep_lines(src)
    start = clean.index(signature_start)
    j, depth = clean.index("{", start), 0
    while True:
        if clean[j] == "{":
            depth += 1
        elif clean[j] == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return src.count("\n", 0, start) + 1, src.count("\n", 0, j) + 1, src[start:j + 1]


def defines(path):
    return [l.rstrip() for l in open(path, encoding="utf-8", errors="replace") if DEFINE_OK.match(l)]


def structs(path, names):
    src = open(path, encoding="utf-8", errors="replace").read()
    out = []
    for nm in names:
        m = re.search(r"struct\s+alignas\(32\)\s+%s\s*\{" % nm, src)
        end = src.find("};", m.start()) + 2
        out.append(src[m.start():end])
    return out


def main():
    if not os.path.isdir(BIND):
        sys.exit(f"ref_extract: {BIND} not found (the reference tree exists only in the build container)")
    os.makedirs(OUT_DIR, exist_ok=True)
    hash_cu = os.path.join(BIND, "field", "Hash3DAnchored_cuda.cu")
    samp_cu = os.path.join(BIND, "PtsSampler", "PersSampler_cuda.cu")
    parts = ["// GENERATED by oracle/ref_extract.py from the reference tree -- build output, never committed.\n"]
    for p in (os.path.join(BIND, "Utils", "Common.h"), os.path.join(BIND, "field", "Hash3DAnchored.h"),
              os.path.join(BIND, "PtsSampler", "PersSampler.h"), samp_cu):
        parts.append(f"// ---- #define constants of {os.path.relpath(p, REF)}")
        parts += defines(p)
    parts.append("// ---- PtsSampler/PersSampler.h structs")
    parts += structs(os.path.join(BIND, "PtsSampler", "PersSampler.h"), ["TransInfo", "TreeNode", "EdgePool"])
    manifest = []
    for p in (hash_cu, samp_cu):
        for first, last, text in device_functions(p):
            parts.append(f"// ---- {os.path.relpath(p, REF)}:{first}-{last}")
            parts.append(text)
            name = re.search(r"(\w+)\s*\(", text[text.find("__"):]).group(1)
            manifest.append(f"{os.path.relpath(p, REF)}:{first}-{last} {name}")
    with open(os.path.join(OUT_DIR, "ref_kernels.inc"), "w") as f:
        f.write("\n".join(parts) + "\n")
    # host member functions of PersOctree that are plain C++ over std::vector<TreeNode> between a torch prologue and
    # epilogue (PtsSampler/PersSampler.cpp): compiled by ref_driver.cpp against a stand-in Tensor (ref_shim/torch_stub.h)
    samp_cpp = os.path.join(BIND, "PtsSampler", "PersSampler.cpp")
    hparts = ["// GENERATED by oracle/ref_extract.py from the reference tree -- build output, never committed.\n"]
    for sig in ("void PersOctree::ProcOctree(", "void PersOctree::ConstructEdgePool("):
        first, last, text = host_function(samp_cpp, sig)
        hparts.append(f"// ---- {os.path.relpath(samp_cpp, REF)}:{first}-{last}")
        hparts.append(text)
        manifest.append(f"{os.path.relpath(samp_cpp, REF)}:{first}-{last} {sig[5:-1]}")
    with open(os.path.join(OUT_DIR, "ref_host_fns.inc"), "w") as f:
        f.write("\n".join(hparts) + "\n")
    # PersOctree CONSTRUCTION (PersSampler.cpp:7-417 and 516-895: DistanceSummary, GetVisiCams, the constructor,
    # ProcOctree, ConstructTreeNode, PCA, ConstructTrans, ConstructEdgePool) is torch tensor code through and through:
    # it is compiled against the REAL libtorch of this image, on the CPU device (ref_driver_torch.cpp), with the
    # reference's own macros (Utils/Common.h) and class declaration (PersSampler.h).
    src = open(samp_cpp, encoding="utf-8", errors="replace").read()
    clean = strip_comments_keep_lines(src)
    a0 = clean.index("using Tensor = torch::Tensor;")
    a1 = clean.index("void PersSampler::VisWarpedPoints(")
    b0 = clean.index("void PersOctree::ConstructTreeNode(")
    b1 = clean.index("PersSampler::PersSampler(")
    b1 = clean.rfind("}", 0, b1) + 1
    hdr = open(os.path.join(BIND, "PtsSampler", "PersSampler.h"), encoding="utf-8", errors="replace").read()
    c0 = hdr.index("class PersOctree {")
    c1 = hdr.index("};", c0) + 2
    common = [l.rstrip() for l in open(os.path.join(BIND, "Utils", "Common.h"), encoding="utf-8", errors="replace")
              if l.startswith("#define") and not l.startswith("#define PRINT_VAL")]
    oparts = ["// GENERATED by oracle/ref_extract.py from the reference tree -- build output, never committed.\n",
              "// ---- gfnerf/bindings/Utils/Common.h #defines"] + common + [
              "// ---- gfnerf/bindings/PtsSampler/PersSampler.h #defines, structs, class PersOctree"] + \
        defines(os.path.join(BIND, "PtsSampler", "PersSampler.h")) + \
        structs(os.path.join(BIND, "PtsSampler", "PersSampler.h"), ["TransInfo", "TreeNode", "EdgePool"]) + [
              hdr[c0:c1],
              f"// ---- gfnerf/bindings/PtsSampler/PersSampler.cpp:{src.count(chr(10), 0, a0) + 1}-{src.count(chr(10), 0, a1)}",
              src[a0:a1],
              f"// ---- gfnerf/bindings/PtsSampler/PersSampler.cpp:{src.count(chr(10), 0, b0) + 1}-{src.count(chr(10), 0, b1) + 1}",
              src[b0:b1]]
    with open(os.path.join(OUT_DIR, "ref_octree_src.inc"), "w") as f:
        f.write("\n".join(oparts) + "\n")
    manifest.append(f"{os.path.relpath(samp_cpp, REF)}:{src.count(chr(10), 0, a0) + 1}-{src.count(chr(10), 0, a1)},"
                    f"{src.count(chr(10), 0, b0) + 1}-{src.count(chr(10), 0, b1) + 1} PersOctree construction (libtorch build)")
    with open(os.path.join(OUT_DIR, "ref_manifest.txt"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    print("\n".join(manifest))


if __name__ == "__main__":
    main()
