"""gf_octree_proc (C++, csrc/octree_host.cu = PersOctree::ProcOctree, PtsSampler/PersSampler.cpp:154-417) against the
numpy restatement `PersOctree.proc_octree` on the same blobs: byte-identical node blobs and statistics.  Host code,
runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.helpers import load_rig


def _octree(rig):
    from gfnerf_b200.persoctree import PersOctree
    oc = PersOctree.__new__(PersOctree)
    oc.load_blobs(rig["tree_nodes"], rig["pers_trans"])
    n = oc.nodes.shape[0]
    oc.weight_stats = np.full(n, 1000, np.int64)
    oc.alpha_stats = np.full(n, 1000, np.int64)
    oc.visit_cnt = np.zeros(n, np.int64)
    return oc


def _cxx(oc, compact, subdivide, brute_force):
    from gfnerf_b200 import _lib
    L = _lib.lib()
    nodes = np.ascontiguousarray(oc.nodes).view(np.uint8).reshape(-1).copy()
    w, a, v = (np.ascontiguousarray(x, np.int64) for x in (oc.weight_stats, oc.alpha_stats, oc.visit_cnt))
    n_out = C.c_int64(0)
    args = (nodes.ctypes.data, nodes.size // 128, w.ctypes.data, a.ctypes.data, v.ctypes.data, int(compact),
            int(subdivide), int(brute_force))
    _lib.check(L.gf_octree_proc(*args, None, None, None, 0, C.byref(n_out)))
    n = n_out.value
    o_nodes, o_w, o_a = np.empty(n * 128, np.uint8), np.empty(n, np.int64), np.empty(n, np.int64)
    _lib.check(L.gf_octree_proc(*args, o_nodes.ctypes.data, o_w.ctypes.data, o_a.ctypes.data, n, C.byref(n_out)))
    return o_nodes, o_w, o_a


@pytest.mark.parametrize("compact,subdivide,brute", [(True, False, False), (True, True, False), (True, True, True),
                                                      (False, True, False), (False, False, False)])
def test_matches_numpy_restatement(compact, subdivide, brute):
    rig = load_rig("rig8")
    rng = np.random.RandomState(5)
    for trial in range(3):
        oc = _octree(rig)
        n = oc.nodes.shape[0]
        # a training history: some leaves voted empty (pruned by MarkInvalidNodes), random statistics, visit counts
        leaves = np.nonzero(oc.nodes["trans_idx"] >= 0)[0]
        dead = rng.choice(leaves, size=int(len(leaves) * (0.2 + 0.3 * trial)), replace=False)
        oc.nodes["trans_idx"][dead] = -1
        oc.weight_stats = rng.randint(-100, 5000, size=n).astype(np.int64)
        oc.alpha_stats = rng.randint(-100, 5000, size=n).astype(np.int64)
        oc.visit_cnt = rng.randint(0, 12, size=n).astype(np.int64)
        if not compact:
            # pruned leaves stay linked from their parents without the compaction pass: the reference's
            # CHECK_GE(node.childs[st], 0) (PersSampler.cpp:315) aborts, and so do both implementations here
            with pytest.raises(RuntimeError):
                _cxx(oc, compact, subdivide, brute)
            with pytest.raises(RuntimeError):
                oc.proc_octree(compact, subdivide, brute)
            oc.proc_octree(True, False, False)     # a compacted tree (no pruned leaf left) goes through
            n = oc.nodes.shape[0]
            oc.visit_cnt = rng.randint(0, 12, size=n).astype(np.int64)
        got_nodes, got_w, got_a = _cxx(oc, compact, subdivide, brute)
        oc.proc_octree(compact, subdivide, brute)
        ref_nodes = np.ascontiguousarray(oc.nodes).view(np.uint8).reshape(-1)
        assert got_nodes.size == ref_nodes.size
        assert np.array_equal(got_nodes, ref_nodes)
        assert np.array_equal(got_w, oc.weight_stats) and np.array_equal(got_a, oc.alpha_stats)
        # structure: parent / child links are mutual, every leaf that survived is valid after a compaction
        nodes = got_nodes.view(oc.nodes.dtype)
        for u in range(nodes.shape[0]):
            for c in nodes["childs"][u]:
                if c >= 0:
                    assert nodes["parent"][c] == u
        if compact and not subdivide:
            assert ((nodes["is_leaf_node"] == 0) | (nodes["trans_idx"] >= 0)).all()


def test_errors():
    from gfnerf_b200 import _lib
    L = _lib.lib()
    n_out = C.c_int64(0)
    assert L.gf_octree_proc(None, 0, None, None, None, 1, 0, 0, None, None, None, 0, C.byref(n_out)) != 0
    assert b"gf_octree_proc" in L.gf_last_error()
    rig = load_rig("rig8")
    oc = _octree(rig)
    nodes = np.ascontiguousarray(oc.nodes).view(np.uint8).reshape(-1).copy()
    w = np.full(oc.nodes.shape[0], 1000, np.int64)
    out = np.empty(128, np.uint8)
    rc = L.gf_octree_proc(nodes.ctypes.data, nodes.size // 128, w.ctypes.data, w.ctypes.data, w.ctypes.data, 1, 1, 1,
                          out.ctypes.data, w.ctypes.data, w.ctypes.data, 1, C.byref(n_out))
    assert rc != 0 and b"capacity" in L.gf_last_error()


@pytest.mark.parametrize("name", ["rig8", "rig20"])
def test_builder_reproduces_the_fixture(name):
    """gf_octree_build (C++, csrc/octree_build.cu = PersOctree::PersOctree + ConstructTreeNode + ConstructTrans,
    PersSampler.cpp:92-152, 516-831) on the 64-camera rig against the committed fixture, which the numpy restatement
    `PersOctree(...)` built from the same cameras and the same (replayed) random draws: the node blob byte for byte,
    the leaf transforms to fp32 rounding -- PCA components up to their sign, which an eigen-decomposition leaves open."""
    from gfnerf_b200.persoctree import TRANS_INFO_DTYPE, aerial_rig, search_order_table
    from gfnerf_b200.perssampler import build_octree
    rig = load_rig(name)          # 64 cameras (parity tests) / 400 cameras (the bench rig, BASELINE config 2)
    c2w, intri, bounds = aerial_rig(n_side=int(rig["n_side"]), extent=float(rig["extent"]), seed=1)
    assert np.array_equal(c2w, rig["c2w"])
    oc = build_octree(16, 512.0, 1.5, c2w, intri, bounds, seed=0)
    assert np.array_equal(oc.tree_nodes_blob(), rig["tree_nodes"])
    assert np.array_equal(oc.search_order, search_order_table())
    got, ref = oc.trans, rig["pers_trans"].view(TRANS_INFO_DTYPE)
    assert got.shape == ref.shape
    assert np.allclose(got["w2xz"], ref["w2xz"], rtol=1e-4, atol=1e-4 * np.abs(ref["w2xz"]).max())
    assert np.array_equal(got["center"], ref["center"]) and np.array_equal(got["side_len"], ref["side_len"])
    assert np.allclose(got["dis_summary"], ref["dis_summary"], rtol=1e-4)
    sign = np.sign((got["weight"] * ref["weight"]).sum(-1, keepdims=True))
    assert (sign != 0).all()
    err = np.abs(got["weight"] * sign - ref["weight"]).max(-1) / np.abs(ref["weight"]).max(-1)
    assert err.max() < 1e-3
    if name == "rig8":
        # the host loops are OpenMP-parallel with fixed-slice sums: the same bytes whatever the thread count
        import subprocess
        import sys
        code = ("import sys, hashlib; sys.path.insert(0, %r); from gfnerf_b200.persoctree import aerial_rig; "
                "from gfnerf_b200.perssampler import build_octree; c, i, b = aerial_rig(n_side=8, extent=4.0, seed=1); "
                "o = build_octree(16, 512.0, 1.5, c, i, b, seed=0); "
                "print(hashlib.sha1(o.pers_trans_blob().tobytes() + o.tree_nodes_blob().tobytes()).hexdigest())")
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        outs = {subprocess.check_output([sys.executable, "-c", code % root], env=dict(os.environ, OMP_NUM_THREADS=t),
                                        text=True).strip().splitlines()[-1] for t in ("1", "5")}
        assert len(outs) == 1, outs
        # errors
        with pytest.raises(RuntimeError):
            build_octree(16, 512.0, 1.5, c2w[:, :2], intri, bounds)


def test_edge_pool_matches_numpy_restatement():
    """gf_octree_edge_pool (PersOctree::ConstructEdgePool, PersSampler.cpp:833-893) against a vectorised numpy
    restatement of the same pair loop: identical 64-byte records in the same order."""
    from gfnerf_b200 import _lib
    rig = load_rig("rig8")
    oc = _octree(rig)
    nodes = np.ascontiguousarray(oc.nodes).view(np.uint8).reshape(-1).copy()
    n = C.c_int64(0)
    L = _lib.lib()
    _lib.check(L.gf_octree_edge_pool(nodes.ctypes.data, oc.nodes.shape[0], None, 0, C.byref(n)))
    pool = np.zeros(n.value * 64, np.uint8)
    _lib.check(L.gf_octree_edge_pool(nodes.ctypes.data, oc.nodes.shape[0], pool.ctypes.data, n.value, C.byref(n)))
    edge_dt = np.dtype({"names": ["a", "b", "center", "dir_0", "dir_1"],
                        "formats": ["<i8", "<i8", ("<f4", 3), ("<f4", 3), ("<f4", 3)],
                        "offsets": [0, 8, 16, 28, 40], "itemsize": 64})
    got = pool.view(edge_dt)
    valid = np.nonzero(oc.nodes["trans_idx"] >= 0)[0]
    c, s, t = oc.nodes["center"][valid], oc.nodes["side_len"][valid], oc.nodes["trans_idx"][valid]
    ref = []
    for ia in range(len(valid)):
        ib = np.arange(ia + 1, len(valid))
        if ib.size == 0:
            continue
        a_small = s[ia] <= s[ib]                                   # u = the smaller of the pair (a on ties)
        cu = np.where(a_small[:, None], c[ia][None], c[ib])
        su = np.where(a_small, s[ia], s[ib])
        cv = np.where(a_small[:, None], c[ib], c[ia][None])
        sv = np.where(a_small, s[ib], s[ia])
        ln = (su * np.float32(.5)).astype(np.float32)
        for j in range(ib.size):
            for axis in range(3):
                for sgn in (1, -1):
                    pt = cu[j].copy()
                    pt[axis] = pt[axis] + ln[j] if sgn > 0 else pt[axis] - ln[j]
                    bias = np.abs((pt - cv[j]) / sv[j] * np.float32(2.)).max()
                    if bias < np.float32(1. + 1e-4):
                        d0, d1 = np.zeros(3, np.float32), np.zeros(3, np.float32)
                        d0[1 if axis == 0 else 0] = ln[j]
                        d1[1 if axis == 2 else 2] = ln[j]
                        ref.append((t[ia], t[ib[j]], pt.astype(np.float32), d0, d1))
    assert len(ref) == got.shape[0] > 0
    assert np.array_equal(got["a"], np.array([r[0] for r in ref])) and np.array_equal(got["b"], np.array([r[1] for r in ref]))
    for k, name in ((2, "center"), (3, "dir_0"), (4, "dir_1")):
        assert np.array_equal(got[name], np.stack([r[k] for r in ref])), name
