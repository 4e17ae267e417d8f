#!/usr/bin/env python
"""Generates tests/golden/ref_kernels.npz by RUNNING THE REFERENCE's own device code -- the __global__ / __device__
function bodies of gfnerf/bindings/field/Hash3DAnchored_cuda.cu and gfnerf/bindings/PtsSampler/PersSampler_cuda.cu,
extracted where they lie under /root/reference and compiled for the host by `make -C oracle ref`
(oracle/ref_extract.py, oracle/ref_driver.cpp, oracle/ref_shim/) -- on fixed inputs.

The fixture holds the inputs too, so nothing is re-derived on another machine.  It is what pins Hash3DAnchored
(rows, blend, gradient), the octree traversal (leaf lists), the march (sample counts, positions), the occupancy
vote and the cold queries of the oracle -- and through the oracle the CUDA kernels -- to the reference's code rather
than to our reading of it.  Outputs are those of the "fma" flavour (g++ contracting mul+add pairs, the analogue of
the nvcc -fmad=true build the reference ships); integer outputs are identical in both flavours.

  make -C oracle ref && python tests/golden/make_golden_ref_kernels.py      # build container only
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_host as rh  # noqa: E402
from tests.helpers import hash_inputs, load_rig  # noqa: E402


def edge_pool_of(tree_nodes):
    """The host builder of the shipped library (gf_octree_edge_pool is host-only code)."""
    from gfnerf_b200 import _lib
    L = _lib.lib()
    nodes = np.ascontiguousarray(tree_nodes, np.uint8)
    n = C.c_int64(0)
    _lib.check(L.gf_octree_edge_pool(nodes.ctypes.data, nodes.size // 128, None, 0, C.byref(n)))
    pool = np.zeros(max(n.value, 1) * 64, np.uint8)
    _lib.check(L.gf_octree_edge_pool(nodes.ctypes.data, nodes.size // 128, pool.ctypes.data, n.value, C.byref(n)))
    return pool[:n.value * 64]


def main(flavour="fma"):
    from gfnerf_b200.persoctree import rig_rays, search_order_table
    fx = {}
    # ---- Hash3DAnchored: power-of-two table with a bias pool, and a non-power-of-two table
    for tag, log2T, local, n_vol, use_bias in (("h0", 10, None, 5, True), ("h1", 10, 48 * 16, 3, False)):
        feat, prim, bias, pts, anchors = hash_inputs(1024, n_vol, log2T, seed=17 + len(tag) + (local or 0))
        if local:
            feat = np.ascontiguousarray(feat[:16 * local])
        L = feat.shape[0] // 16
        if use_bias:
            bias = np.random.RandomState(2).uniform(100, 1100, size=bias.shape).astype(np.float32)  # Hash3DAnchored.cpp:58
        pts[:4] = [[0, 0, 0], [1, 1, 1], [0.5, 0.25, 0.125], [0.999999, 0, 1]]          # lattice nodes / domain corners
        g = (np.random.RandomState(3).normal(size=(1024, 32)) * 1e-3).astype(np.float32)
        g[::5] = 0                                                                     # rows the `if (w0 != 0 || w1 != 0)` skips
        fx.update({f"{tag}_feat": feat, f"{tag}_prim": prim, f"{tag}_bias": bias, f"{tag}_pts": pts,
                   f"{tag}_anchors": anchors, f"{tag}_grad": g, f"{tag}_local": np.int64(L),
                   f"{tag}_out": rh.hash_forward(feat, prim, bias, pts, anchors, flavour).astype(np.float16),
                   f"{tag}_gtable": rh.hash_backward(L, prim, bias, pts, anchors, g, flavour)})
        # a gradient hitting every table row at most once: the fp16 atomics accumulate nothing, the result is exact
        few = slice(0, 6)
        fx[f"{tag}_gtable_few"] = rh.hash_backward(L, prim, bias, pts[few], anchors[few], g[few] * 50, flavour)
    # ---- PersSampler on the prebuilt rig (tests/golden/rig8.npz)
    rig = load_rig("rig8")
    so = search_order_table().reshape(-1).astype(np.uint8)
    R = 96
    o, d, _ = rig_rays(rig["c2w"], rig["intri"], R, seed=5)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    o[-2:] = [[0, 0, 900.0], [5000.0, 0, 2.0]]                                         # rays that miss the tree
    d[-2:] = [[0, 0, 1.0], [1.0, 0, 0]]
    d[-3] = [0.0, 0.0, -1.0]                                                           # axis-parallel: the |d| < 1e-6 slab case
    for mode, noise in (("eval", np.ones(1024 + R + 10, np.float32)),
                        ("train", np.random.RandomState(0).uniform(0.5, 1.5, 1024 + R + 10).astype(np.float32) * 4.0)):
        s = rh.get_samples(o, d, noise, rig["tree_nodes"], rig["pers_trans"], so, flavour=flavour)
        m = np.arange(1024)[None, :] < s["counts"][:, None]
        max_oct = int(s["n_oct"].max())
        fx.update({f"s_{mode}_noise": noise, f"s_{mode}_counts": s["counts"], f"s_{mode}_n_oct": s["n_oct"].astype(np.int32),
                   f"s_{mode}_oct_idx": s["oct_idx"][:, :max_oct].astype(np.int32),
                   f"s_{mode}_oct_nf": s["oct_nf"][:, :max_oct], f"s_{mode}_first_oct_dis": s["first_oct_dis"],
                   f"s_{mode}_pts_idx_start_end": s["pts_idx_start_end"],
                   f"s_{mode}_anchors": s["anchors"][m].astype(np.int32), f"s_{mode}_ts": s["ts"][m],
                   f"s_{mode}_dists": s["dists"][m], f"s_{mode}_warp_pts": s["warp_pts"][m],
                   f"s_{mode}_dirs_ok": np.bool_(
                       np.array_equal(s["dirs"][m], np.broadcast_to(d[:, None, :], s["dirs"].shape)[m]))})
        assert not s["anchors"][~m].any() and not s["ts"][~m].any()                    # padding stays zero
        if mode == "eval":
            ev = s
    fx.update({"s_rays_o": o, "s_rays_d": d, "s_search_order": so})
    # ---- UpdateOctNodes (vote + stats + MarkInvalidNodes) on the eval-mode samples
    rng = np.random.RandomState(1)
    w = ((rng.rand(R, 1024) ** 6) * 0.05).astype(np.float32)
    a = ((rng.rand(R, 1024) ** 6) * 0.1).astype(np.float32)
    nodes = rig["tree_nodes"].copy()
    n_nodes = nodes.size // 128
    ws, as_, vc = np.full(n_nodes, 1000, np.int64), np.full(n_nodes, 1000, np.int64), np.zeros(n_nodes, np.int64)
    ws[::7] = 0
    as_[3::11] = -50
    fx.update({"v_weights": w.astype(np.float16), "v_alphas": a.astype(np.float16), "v_ws_in": ws.copy(),
               "v_as_in": as_.copy()})
    w, a = fx["v_weights"].astype(np.float32), fx["v_alphas"].astype(np.float32)       # what the test will feed
    rh.update_oct_nodes(ev["pts_idx_start_end"], ev["anchors"][..., 1].reshape(-1), w.reshape(-1), a.reshape(-1),
                        nodes, ws, as_, vc, flavour=flavour)
    fx.update({"v_ws": ws, "v_as": as_, "v_cnt": vc,
               "v_trans_idx": nodes.view(np.int64).reshape(-1, 16)[:, 12].copy()})
    # ---- cold queries
    anchors = rng.randint(-2, n_nodes + 2, size=2000).astype(np.int64)
    pts = rng.uniform(-3, 3, size=(2000, 3)).astype(np.float32)
    fx.update({"q_anchors": anchors, "q_pts": pts,
               "q_out": rh.trans_query_frame(rig["tree_nodes"], rig["pers_trans"], anchors, pts, flavour)})
    t_cur = np.sort(rng.uniform(0.05, 12.0, size=(64, 32)).astype(np.float32), axis=1)
    assert R >= 64
    fx.update({"p_t_cur": t_cur, "p_anchors": rh.points_anchors(o[:64], d[:64], t_cur, rig["tree_nodes"], flavour)})
    pool = edge_pool_of(rig["tree_nodes"])
    assert rh.lib(flavour).ref_sizeof_edge_pool() == 64
    eidx = rng.randint(0, pool.size // 64, size=300).astype(np.int64)
    ecoord = rng.uniform(-1, 1, size=(300, 2)).astype(np.float32)
    epts, eids = rh.edge_samples(pool, rig["pers_trans"], eidx, ecoord, flavour)
    fx.update({"e_idx": eidx, "e_coord": ecoord, "e_pts": epts, "e_ids": eids})
    # ---- PersOctree::ProcOctree / ConstructEdgePool (host functions): digests of the reference's outputs
    import hashlib
    from tests.test_ref_kernels import _history
    hrng = np.random.RandomState(5)
    for trial in range(3):
        nodes, w, a, v = _history(rig, trial, hrng)
        for tag, flags in (("c", (True, False, False)), ("cs", (True, True, False)), ("csb", (True, True, True))):
            out = rh.proc_octree(nodes, w, a, v, *flags, flavour=flavour)
            fx[f"o_{trial}_{tag}"] = np.array([out[0].size // 128] + [int(hashlib.sha256(x.tobytes()).hexdigest()[:15], 16)
                                                                      for x in out], np.int64)
    ep = rh.construct_edge_pool(rig["tree_nodes"], flavour)
    used = np.arange(ep.size) % 64 < 52
    fx["o_edge_pool"] = np.array([ep.size // 64, int(hashlib.sha256(ep[used].tobytes()).hexdigest()[:15], 16)], np.int64)
    # ---- PersOctree construction by the reference's own constructor (real libtorch on the CPU; ~40 s)
    if rh.octree_available():
        from gfnerf_b200.persoctree import TRANS_INFO_DTYPE
        from tests.test_ref_octree import _w2c, node_digest
        nodes, trans, so = rh.build_octree(16, 512.0, 1.5, rig["c2w"], _w2c(rig["c2w"]), rig["intri"], rig["bounds"], seed=0)
        t = trans.view(TRANS_INFO_DTYPE)
        fx.update({"t_nodes": np.array([nodes.size // 128, node_digest(nodes)], np.int64), "t_search_order": so,
                   "t_center": t["center"].copy(), "t_side_len": t["side_len"].copy(),
                   "t_dis_summary": t["dis_summary"].copy()})
    path = os.path.join(HERE, "ref_kernels.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB;", rh.lib(flavour).ref_build_flavour().decode())
    print("samples eval/train:", int(fx["s_eval_counts"].sum()), int(fx["s_train_counts"].sum()),
          "leaves/ray max", int(fx["s_eval_n_oct"].max()), "pruned by the vote:",
          int((fx["v_trans_idx"] != rig["tree_nodes"].view(np.int64).reshape(-1, 16)[:, 12]).sum()))


if __name__ == "__main__":
    main()
