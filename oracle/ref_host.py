"""ctypes wrapper of oracle/_ref/libgf_ref_host*.so -- the REFERENCE's own kernel bodies (extracted at build time from
/root/reference by oracle/ref_extract.py) compiled for the host by oracle/ref_driver.cpp.  TEST INFRASTRUCTURE ONLY.

Two flavours are built (oracle/Makefile target `ref`): "off" = -ffp-contract=off (every fp32 operation rounded on its
own) and "fma" = -ffp-contract=fast -mfma (g++ fuses mul+add pairs where it likes, as nvcc -fmad=true does where IT
likes).  The shared objects are git-ignored but travel to the GPU box with the snapshot, like every built .so.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_libs = {}

_vp, _i64, _f32, _int = C.c_void_p, C.c_int64, C.c_float, C.c_int


def path(flavour="off"):
    return os.path.join(REF_DIR, "libgf_ref_host.so" if flavour == "off" else "libgf_ref_host_fma.so")


def available(flavour="off"):
    return os.path.exists(path(flavour))


def lib(flavour="off"):
    if flavour not in _libs:
        L = C.CDLL(path(flavour))
        L.ref_build_flavour.restype = C.c_char_p
        L.ref_sizeof_edge_pool.restype = C.c_int64
        _libs[flavour] = L
    return _libs[flavour]


def set_threads(n, flavour="off"):
    """Host threads a kernel launch of this flavour's library is spread over.  1 (the default) = one thread instance
    after the other, deterministic -- what the tests use; os.cpu_count() = the reference's kernels on all host cores
    (bench.py's CPU arm).  Outputs that do not depend on atomic ORDER (everything except the fp16 gradient sums and
    where in the leaf pool a ray's list lands) are identical either way."""
    lib(flavour).ref_set_threads(_int(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def hash_forward(feat_f32, prim_pool, bias_pool, pts, anchors, flavour="off"):
    """Hash3DAnchoredFunction::forward: table cast to fp16, kernel, output widened to fp32.  -> [n,32] fp32"""
    n, n_vol = pts.shape[0], prim_pool.shape[1]
    local = feat_f32.shape[0] // 16
    table = _c(feat_f32, np.float32).astype(np.float16)
    prim, bias = _c(prim_pool, np.int32), _c(bias_pool, np.float32)
    fidx = (np.arange(16) * local).astype(np.int32)
    fsize = np.full(16, local, np.int32)
    out = np.zeros((n, 32), np.float16)
    pts, anchors = _c(pts, np.float32), _c(anchors, np.int64)
    lib(flavour).ref_hash_forward(_int(n), _int(n_vol), _p(table), _p(prim), _p(fidx), _p(fsize), _p(bias), _p(pts),
                                  _p(anchors), _p(out))
    return out.astype(np.float32)


def hash_backward(local_size, prim_pool, bias_pool, pts, anchors, grad_out, flavour="off"):
    """Hash3DAnchoredFunction::backward: (grad * 128) -> fp16, kernel with fp16 half2 atomics (serial order), result
    widened and divided by 128.  -> [16*local_size, 2] fp32"""
    n, n_vol = pts.shape[0], prim_pool.shape[1]
    prim, bias = _c(prim_pool, np.int32), _c(bias_pool, np.float32)
    fidx = (np.arange(16) * local_size).astype(np.int32)
    fsize = np.full(16, local_size, np.int32)
    gin = (_c(grad_out, np.float32) * np.float32(128.0)).astype(np.float16)
    gout = np.zeros((16 * local_size, 2), np.float16)
    pts, anchors = _c(pts, np.float32), _c(anchors, np.int64)
    lib(flavour).ref_hash_backward(_int(n), _int(n_vol), _p(prim), _p(fidx), _p(fsize), _p(bias), _p(pts),
                                   _p(anchors), _p(gin), _p(gout))
    return gout.astype(np.float32) / np.float32(128.0)


def get_samples(rays_o, rays_d_unit, noise, tree_nodes, pers_trans, search_order, global_near=0.01,
                sample_l=1.0 / 256, scale_by_dis=True, max_oct=1024, flavour="off"):
    R, S = rays_o.shape[0], 1024
    rays_o, rays_d_unit, noise = _c(rays_o, np.float32), _c(rays_d_unit, np.float32), _c(noise, np.float32)
    nodes, trans = _c(tree_nodes, np.uint8).copy(), _c(pers_trans, np.uint8).copy()
    so = _c(search_order, np.uint8)
    out = dict(world_pts=np.zeros((R, S, 3), np.float32), warp_pts=np.zeros((R, S, 3), np.float32),
               dirs=np.zeros((R, S, 3), np.float32), dists=np.zeros((R, S), np.float32),
               ts=np.zeros((R, S), np.float32), anchors=np.zeros((R, S, 3), np.int64),
               pts_idx_start_end=np.zeros((R, 2), np.int64), first_oct_dis=np.zeros(R, np.float32),
               n_oct=np.zeros(R, np.int64), oct_idx=np.full((R, max_oct), -1, np.int64),
               oct_nf=np.zeros((R, max_oct, 2), np.float32))
    lib(flavour).ref_get_samples(_i64(R), _p(rays_o), _p(rays_d_unit), _p(noise), _p(nodes), _p(trans), _p(so),
                                 _f32(global_near), _f32(sample_l), _int(int(scale_by_dis)), _i64(max_oct),
                                 _p(out["world_pts"]), _p(out["warp_pts"]), _p(out["dirs"]), _p(out["dists"]),
                                 _p(out["ts"]), _p(out["anchors"]), _p(out["pts_idx_start_end"]),
                                 _p(out["first_oct_dis"]), _p(out["n_oct"]), _p(out["oct_idx"]), _p(out["oct_nf"]))
    se = out["pts_idx_start_end"]
    out["counts"] = (se[:, 1] - se[:, 0]).astype(np.int32)
    return out


def update_oct_nodes(pts_idx_start_end, oct_indices, weights, alphas, tree_nodes, weight_stats, alpha_stats, visit_cnt,
                     flavour="off"):
    """In place on tree_nodes (uint8 blob) and the three int64 stat arrays."""
    R = pts_idx_start_end.shape[0]
    lib(flavour).ref_update_oct_nodes(_i64(R), _p(_c(pts_idx_start_end, np.int64)), _p(_c(oct_indices, np.int64)),
                                      _p(_c(weights, np.float32)), _p(_c(alphas, np.float32)), _p(tree_nodes),
                                      _i64(tree_nodes.size // 128), _p(weight_stats), _p(alpha_stats), _p(visit_cnt))


def trans_query_frame(tree_nodes, pers_trans, anchors, world_pts, flavour="off"):
    n = world_pts.shape[0]
    nodes, trans = _c(tree_nodes, np.uint8), _c(pers_trans, np.uint8)
    out = np.zeros((n, 3), np.float32)
    lib(flavour).ref_trans_query_frame(_i64(n), _p(nodes), _i64(nodes.size // 128), _p(trans),
                                       _p(_c(anchors, np.int64)), _p(_c(world_pts, np.float32)), _p(out))
    return out


def points_anchors(rays_o, rays_d, t_cur, tree_nodes, flavour="off"):
    t_cur = _c(t_cur, np.float32)
    R, S = t_cur.shape[0], t_cur.shape[1]
    nodes = _c(tree_nodes, np.uint8)
    out = np.full((R, S), -1, np.int64)
    lib(flavour).ref_points_anchors(_i64(R), _i64(S), _p(_c(rays_o, np.float32)), _p(_c(rays_d, np.float32)),
                                    _p(t_cur), _p(nodes), _i64(nodes.size // 128), _p(out))
    return out


def edge_samples(edge_pool, pers_trans, edge_idx, edge_coords, flavour="off"):
    edge_idx = _c(edge_idx, np.int64)
    n = edge_idx.shape[0]
    pts, idx = np.zeros((n, 2, 3), np.float32), np.zeros((n, 2), np.int64)
    lib(flavour).ref_edge_samples(_i64(n), _p(_c(edge_pool, np.uint8)), _p(_c(pers_trans, np.uint8)), _p(edge_idx),
                                  _p(_c(edge_coords, np.float32)), _p(pts), _p(idx))
    return pts, idx


def mark_invisible_nodes(tree_nodes, intri, w2c, bounds, flavour="off"):
    """In place on tree_nodes."""
    lib(flavour).ref_mark_invisible_nodes(_i64(tree_nodes.size // 128), _i64(intri.shape[0]), _p(tree_nodes),
                                          _p(_c(intri, np.float32)), _p(_c(w2c, np.float32)),
                                          _p(_c(bounds, np.float32)))


def set_block_idxs(tree_nodes, centers, flavour="off"):
    """In place on tree_nodes."""
    centers = _c(centers, np.float32)
    lib(flavour).ref_set_block_idxs(_i64(tree_nodes.size // 128), _i64(centers.shape[0]), _p(tree_nodes), _p(centers))


# ---------------------------------------------------------------- the same bodies compiled by nvcc (GPU box only)
CUDA_LIB = os.path.join(REF_DIR, "libgf_ref_cuda.so")
_cuda_lib = None


def cuda_available():
    return os.path.exists(CUDA_LIB)


def cuda_lib():
    """oracle/_ref/libgf_ref_cuda.so (`make -C oracle ref_cuda`): the reference's device functions as real CUDA for
    sm_100a.  Entry points take torch CUDA tensors' data_ptr()s; see oracle/ref_driver_cuda.cu."""
    global _cuda_lib
    if _cuda_lib is None:
        _cuda_lib = C.CDLL(CUDA_LIB)
    return _cuda_lib


def _dp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def cuda_hash_forward(feat_f32, prim_pool, bias_pool, pts, anchors):
    """torch CUDA tensors in, fp32 [n,32] out (Hash3DAnchoredFunction::forward with the reference's casts)."""
    import torch
    n, n_vol, local = pts.shape[0], prim_pool.shape[1], feat_f32.shape[0] // 16
    dev = pts.device
    table = feat_f32.to(torch.float16).contiguous()
    fidx = (torch.arange(16, device=dev) * local).to(torch.int32)
    fsize = torch.full((16,), local, dtype=torch.int32, device=dev)
    out = torch.zeros((n, 32), dtype=torch.float16, device=dev)
    rc = cuda_lib().refcu_hash_forward(_int(n), _int(n_vol), _dp(table), _dp(prim_pool.int().contiguous()), _dp(fidx),
                                       _dp(fsize), _dp(bias_pool.float().contiguous()), _dp(pts.float().contiguous()),
                                       _dp(anchors.long().contiguous()), _dp(out))
    assert rc == 0, f"refcu_hash_forward: cudaError {rc}"
    return out.float()


def cuda_hash_backward(local_size, prim_pool, bias_pool, pts, anchors, grad_out):
    import torch
    n, n_vol, dev = pts.shape[0], prim_pool.shape[1], pts.device
    fidx = (torch.arange(16, device=dev) * local_size).to(torch.int32)
    fsize = torch.full((16,), local_size, dtype=torch.int32, device=dev)
    gin = (grad_out.float() * 128.0).to(torch.float16).contiguous()
    gout = torch.zeros((16 * local_size, 2), dtype=torch.float16, device=dev)
    rc = cuda_lib().refcu_hash_backward(_int(n), _int(n_vol), _dp(prim_pool.int().contiguous()), _dp(fidx), _dp(fsize),
                                        _dp(bias_pool.float().contiguous()), _dp(pts.float().contiguous()),
                                        _dp(anchors.long().contiguous()), _dp(gin), _dp(gout))
    assert rc == 0, f"refcu_hash_backward: cudaError {rc}"
    return gout.float() / 128.0


def cuda_get_samples(rays_o, rays_d_unit, noise, tree_nodes, pers_trans, search_order, global_near=0.01,
                     sample_l=1.0 / 256, scale_by_dis=True, max_oct=1024):
    """torch CUDA tensors in; dict of torch CUDA tensors out (dense reference layout)."""
    import torch
    R, S, dev = rays_o.shape[0], 1024, rays_o.device
    z = lambda *sh, dt=torch.float32: torch.zeros(sh, dtype=dt, device=dev)
    out = dict(world_pts=z(R, S, 3), warp_pts=z(R, S, 3), dirs=z(R, S, 3), dists=z(R, S), ts=z(R, S),
               anchors=z(R, S, 3, dt=torch.int64), pts_idx_start_end=z(R, 2, dt=torch.int64), first_oct_dis=z(R),
               oct_idx_start_end=z(R, 2, dt=torch.int64))
    rc = cuda_lib().refcu_get_samples(
        _i64(R), _dp(rays_o.float().contiguous()), _dp(rays_d_unit.float().contiguous()), _dp(noise.float().contiguous()),
        _dp(tree_nodes), _dp(pers_trans), _dp(search_order), _f32(global_near), _f32(sample_l),
        _int(int(scale_by_dis)), _i64(max_oct), _dp(out["world_pts"]), _dp(out["warp_pts"]), _dp(out["dirs"]),
        _dp(out["dists"]), _dp(out["ts"]), _dp(out["anchors"]), _dp(out["pts_idx_start_end"]),
        _dp(out["first_oct_dis"]), _dp(out["oct_idx_start_end"]))
    assert rc == 0, f"refcu_get_samples: cudaError {rc}"
    se = out["pts_idx_start_end"]
    out["counts"] = (se[:, 1] - se[:, 0]).to(torch.int32)
    return out


def cuda_mark_invisible_nodes(tree_nodes, intri, w2c, bounds):
    """torch CUDA tensors; in place on tree_nodes (uint8 [128 * n])."""
    rc = cuda_lib().refcu_mark_invisible_nodes(_i64(tree_nodes.numel() // 128), _i64(intri.shape[0]), _dp(tree_nodes),
                                               _dp(intri.float().contiguous()), _dp(w2c.float().contiguous()),
                                               _dp(bounds.float().contiguous()))
    assert rc == 0, f"refcu_mark_invisible_nodes: cudaError {rc}"


def cuda_set_block_idxs(tree_nodes, centers):
    """torch CUDA tensors; in place on tree_nodes."""
    centers = centers.float().contiguous()
    rc = cuda_lib().refcu_set_block_idxs(_i64(tree_nodes.numel() // 128), _i64(centers.shape[0]), _dp(tree_nodes),
                                         _dp(centers))
    assert rc == 0, f"refcu_set_block_idxs: cudaError {rc}"


# ---------------------------------------------------------------- host member functions of PersOctree
def proc_octree(tree_nodes, weight_stats, alpha_stats, visit_cnt, compact, subdivide, brute_force, flavour="off"):
    """PersOctree::ProcOctree (PtsSampler/PersSampler.cpp:154-417), the reference's own body.  -> (nodes blob uint8,
    weight_stats, alpha_stats) of the processed tree."""
    L = lib(flavour)
    L.ref_proc_octree.restype = C.c_int64
    nodes = _c(tree_nodes, np.uint8)
    n_in = nodes.size // 128
    w, a, v = _c(weight_stats, np.int64), _c(alpha_stats, np.int64), _c(visit_cnt, np.int64)
    args = (_p(nodes), _i64(n_in), _p(w), _p(a), _p(v), _int(int(compact)), _int(int(subdivide)), _int(int(brute_force)))
    n = L.ref_proc_octree(*args, None, None, None, _i64(0))
    if n < 0:
        raise RuntimeError(f"a CHECK of the reference's ProcOctree failed (ref_host_fns.inc line {-n})")
    out, wo, ao = np.zeros(n * 128, np.uint8), np.zeros(n, np.int64), np.zeros(n, np.int64)
    assert L.ref_proc_octree(*args, _p(out), _p(wo), _p(ao), _i64(n)) == n
    return out, wo, ao


def construct_edge_pool(tree_nodes, flavour="off"):
    """PersOctree::ConstructEdgePool (PersSampler.cpp:833-895).  -> uint8 blob of 64-byte EdgePool entries."""
    L = lib(flavour)
    L.ref_construct_edge_pool.restype = C.c_int64
    nodes = _c(tree_nodes, np.uint8)
    n = L.ref_construct_edge_pool(_p(nodes), _i64(nodes.size // 128), None, _i64(0))
    pool = np.zeros(max(n, 1) * 64, np.uint8)
    assert L.ref_construct_edge_pool(_p(nodes), _i64(nodes.size // 128), _p(pool), _i64(n)) == n
    return pool[:n * 64]


# ---------------------------------------------------------------- PersOctree construction against real libtorch (CPU)
OCTREE_LIB = os.path.join(REF_DIR, "libgf_ref_octree.so")
_octree_lib = None


def octree_available():
    return os.path.exists(OCTREE_LIB)


def octree_lib():
    """oracle/_ref/libgf_ref_octree.so (`make -C oracle ref_octree`): PtsSampler/PersSampler.cpp:9-417, 516-895 compiled
    against this image's libtorch with kCUDA -> kCPU (oracle/ref_driver_torch.cpp)."""
    global _octree_lib
    if _octree_lib is None:
        import torch  # noqa: F401  (loads libtorch / libc10 into the process first)
        _octree_lib = C.CDLL(OCTREE_LIB)
        _octree_lib.refoct_build.restype = C.c_int64
    return _octree_lib


def build_octree(max_depth, bbox_side_len, split_dist_thres, c2w, w2c, intri, bound, seed=0):
    """The reference's PersOctree constructor.  -> (tree_nodes uint8 blob, pers_trans uint8 blob, search_order u8[64])"""
    L = octree_lib()
    c2w, w2c, intri, bound = (_c(x, np.float32) for x in (c2w, w2c, intri, bound))
    n_trans = C.c_int64(0)
    args = (_i64(max_depth), _f32(bbox_side_len), _f32(split_dist_thres), _p(c2w), _p(w2c), _p(intri), _p(bound),
            _i64(c2w.shape[0]), C.c_uint64(seed))
    n = L.refoct_build(*args, None, None, None, C.byref(n_trans))
    if n < 0:
        raise RuntimeError("the reference's PersOctree constructor failed (see stderr)")
    nodes, trans, so = np.zeros(n * 128, np.uint8), np.zeros(n_trans.value * 576, np.uint8), np.zeros(64, np.uint8)
    assert L.refoct_build(*args, _p(nodes), _p(trans), _p(so), C.byref(n_trans)) == n
    return nodes, trans, so


def construct_trans(rand_pts, c2w, intri0, center, seed=0):
    """The reference's PersOctree::ConstructTrans.  -> 576-byte TransInfo blob (side_len = 0)."""
    rand_pts, c2w = _c(rand_pts, np.float32), _c(c2w, np.float32)
    out = np.zeros(576, np.uint8)
    rc = octree_lib().refoct_construct_trans(_p(rand_pts), _i64(rand_pts.shape[0]), _p(c2w), _i64(c2w.shape[0]),
                                             _p(_c(intri0, np.float32)), _p(_c(center, np.float32)), C.c_uint64(seed),
                                             _p(out))
    if rc != 0:
        raise RuntimeError("the reference's ConstructTrans failed (see stderr)")
    return out
