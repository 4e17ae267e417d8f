"""ctypes front-end of the CPU oracle (oracle/gf_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of gf_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference); never by
the product package.  All arguments are numpy arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgf_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgf_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_charbonnier.restype = C.c_double
        _lib.orc_mlp_param_count.restype = C.c_int64
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


# ---------------------------------------------------------------- hash
def hash_level_scales():
    s = np.zeros(16, np.float32)
    lib().orc_hash_level_scales(_p(s))
    return s


def hash_forward(feat_f32, prim_pool, bias_pool, pts, anchors, scales=None, want_idx=False):
    """-> out [n,32] fp32 (fp16-rounded values) [, idx int32 [n,16,8]]"""
    n = pts.shape[0]
    n_vol = prim_pool.shape[1]
    local = feat_f32.shape[0] // 16
    feat_f32, bias_pool, pts = _f32(feat_f32), _f32(bias_pool), _f32(pts)
    prim_pool = np.ascontiguousarray(prim_pool, np.int32)
    anchors = np.ascontiguousarray(anchors, np.int64)
    scales = hash_level_scales() if scales is None else _f32(scales)
    out = np.zeros((n, 32), np.float32)
    idx = np.zeros((n, 16, 8), np.int32) if want_idx else None
    lib().orc_hash_forward(C.c_int64(n), C.c_int32(n_vol), C.c_int64(local), _p(feat_f32), _p(prim_pool),
                           _p(bias_pool), _p(scales), _p(pts), _p(anchors), _p(out), _p(idx))
    return (out, idx) if want_idx else out


def hash_backward(local_size, prim_pool, bias_pool, pts, anchors, grad_out, scales=None):
    """-> grad_table fp64 [16*local_size, 2]"""
    n = pts.shape[0]
    n_vol = prim_pool.shape[1]
    bias_pool, pts, grad_out = _f32(bias_pool), _f32(pts), _f32(grad_out)
    prim_pool = np.ascontiguousarray(prim_pool, np.int32)
    anchors = np.ascontiguousarray(anchors, np.int64)
    scales = hash_level_scales() if scales is None else _f32(scales)
    g = np.zeros((16 * local_size, 2), np.float64)
    lib().orc_hash_backward(C.c_int64(n), C.c_int32(n_vol), C.c_int64(local_size), _p(prim_pool), _p(bias_pool),
                            _p(scales), _p(pts), _p(anchors), _p(grad_out), _p(g))
    return g


# ---------------------------------------------------------------- sampler
def search_order():
    o = np.zeros(64, np.uint8)
    lib().orc_search_order(_p(o))
    return o


def sampler_get_samples(rays_o, rays_d_unit, noise, tree_nodes, pers_trans, global_near=0.01, sample_l=1.0 / 256,
                        scale_by_dis=True, max_oct=1024, want_oct=False):
    """Dense reference layout.  Returns a dict of numpy arrays."""
    R = rays_o.shape[0]
    S = 1024
    rays_o, rays_d_unit, noise = _f32(rays_o), _f32(rays_d_unit), _f32(noise)
    assert noise.shape[0] >= S + R
    tree_nodes = np.ascontiguousarray(tree_nodes, np.uint8)
    pers_trans = np.ascontiguousarray(pers_trans, np.uint8)
    so = search_order()
    out = dict(
        world_pts=np.zeros((R, S, 3), np.float32), warp_pts=np.zeros((R, S, 3), np.float32),
        dirs=np.zeros((R, S, 3), np.float32), dists=np.zeros((R, S), np.float32), ts=np.zeros((R, S), np.float32),
        anchors=np.zeros((R, S, 3), np.int64), counts=np.zeros(R, np.int32),
        first_oct_dis=np.zeros(R, np.float32), n_oct=np.zeros(R, np.int32))
    if want_oct:
        out["oct_idx"] = np.full((R, max_oct), -1, np.int64)
        out["oct_nf"] = np.zeros((R, max_oct, 2), np.float32)
    lib().orc_sampler_get_samples(
        C.c_int64(R), _p(rays_o), _p(rays_d_unit), _p(noise), _p(tree_nodes), _p(pers_trans), _p(so),
        C.c_float(global_near), C.c_float(sample_l), C.c_int(int(scale_by_dis)), C.c_int64(max_oct),
        _p(out["world_pts"]), _p(out["warp_pts"]), _p(out["dirs"]), _p(out["dists"]), _p(out["ts"]),
        _p(out["anchors"]), _p(out["counts"]), _p(out["first_oct_dis"]), _p(out["n_oct"]),
        _p(out.get("oct_idx")), _p(out.get("oct_nf")))
    return out


def trans_query_frame(tree_nodes, pers_trans, anchors, world_pts):
    n = world_pts.shape[0]
    tree_nodes = np.ascontiguousarray(tree_nodes, np.uint8)
    pers_trans = np.ascontiguousarray(pers_trans, np.uint8)
    out = np.zeros((n, 3), np.float32)
    lib().orc_trans_query_frame(C.c_int64(n), _p(tree_nodes), C.c_int64(tree_nodes.size // 128), _p(pers_trans),
                                _p(np.ascontiguousarray(anchors, np.int64)), _p(_f32(world_pts)), _p(out))
    return out


def update_oct_nodes(counts, oct_indices, weights, alphas, tree_nodes, weight_stats, alpha_stats, visit_cnt):
    """In-place on tree_nodes (uint8 blob), weight_stats, alpha_stats, visit_cnt (int64)."""
    R = counts.shape[0]
    n_nodes = tree_nodes.size // 128
    lib().orc_update_oct_nodes(C.c_int64(R), _p(np.ascontiguousarray(counts, np.int32)),
                               _p(np.ascontiguousarray(oct_indices, np.int64)), _p(_f32(weights)), _p(_f32(alphas)),
                               _p(tree_nodes), C.c_int64(n_nodes), _p(weight_stats), _p(alpha_stats), _p(visit_cnt))


# ---------------------------------------------------------------- composite
def composite_forward(offsets, sigma, delta, rgb, t):
    R = offsets.shape[0] - 1
    V = sigma.shape[0]
    offsets = np.ascontiguousarray(offsets, np.int32)
    out = dict(weights=np.zeros(V, np.float32), alphas=np.zeros(V, np.float32), trans=np.zeros(V, np.float32),
               rgb=np.zeros((R, 3), np.float32), depth=np.zeros(R, np.float32), acc=np.zeros(R, np.float32))
    lib().orc_composite_forward(C.c_int64(R), _p(offsets), _p(_f32(sigma)), _p(_f32(delta)), _p(_f32(rgb)),
                                _p(_f32(t)), _p(out["weights"]), _p(out["alphas"]), _p(out["trans"]),
                                _p(out["rgb"]), _p(out["depth"]), _p(out["acc"]))
    return out


def composite_backward(offsets, sigma, delta, rgb, g_rgb, g_acc=None):
    R = offsets.shape[0] - 1
    V = sigma.shape[0]
    offsets = np.ascontiguousarray(offsets, np.int32)
    d_sigma = np.zeros(V, np.float32)
    d_rgb = np.zeros((V, 3), np.float32)
    lib().orc_composite_backward(C.c_int64(R), _p(offsets), _p(_f32(sigma)), _p(_f32(delta)), _p(_f32(rgb)),
                                 _p(_f32(g_rgb)), _p(None if g_acc is None else _f32(g_acc)), _p(d_sigma), _p(d_rgb))
    return d_sigma, d_rgb


# ---------------------------------------------------------------- MLP
def sh4(dirs_unit):
    dirs_unit = _f32(dirs_unit)
    out = np.zeros((dirs_unit.shape[0], 16), np.float32)
    for i in range(dirs_unit.shape[0]):
        lib().orc_sh4(_p(dirs_unit[i]), _p(out[i]))
    return out


def mlp_param_count(H=64):
    return int(lib().orc_mlp_param_count(C.c_int(H)))


def mlp_forward(params, feat, ray_id, ray_dirs, ray_emb=None, H=64):
    n = feat.shape[0]
    sigma = np.zeros(n, np.float32)
    rgb = np.zeros((n, 3), np.float32)
    lib().orc_mlp_forward(C.c_int64(n), C.c_int(H), _p(_f32(params)), _p(_f32(feat)),
                          _p(np.ascontiguousarray(ray_id, np.int32)), _p(_f32(ray_dirs)),
                          _p(None if ray_emb is None else _f32(ray_emb)), _p(sigma), _p(rgb))
    return sigma, rgb


def mlp_backward(params, feat, ray_id, ray_dirs, ray_emb, d_sigma, d_rgb, H=64):
    n = feat.shape[0]
    R = ray_dirs.shape[0]
    d_feat = np.zeros((n, 32), np.float32)
    d_params = np.zeros(mlp_param_count(H), np.float64)
    d_emb = None if ray_emb is None else np.zeros((R, 32), np.float64)
    lib().orc_mlp_backward(C.c_int64(n), C.c_int(H), _p(_f32(params)), _p(_f32(feat)),
                           _p(np.ascontiguousarray(ray_id, np.int32)), _p(_f32(ray_dirs)),
                           _p(None if ray_emb is None else _f32(ray_emb)), _p(_f32(d_sigma)), _p(_f32(d_rgb)),
                           _p(d_feat), _p(d_params), _p(d_emb))
    return d_feat, d_params, d_emb


# ---------------------------------------------------------------- loss / Adam
def charbonnier(rgb, target, eps=1e-6):
    rgb, target = _f32(rgb), _f32(target)
    g = np.zeros_like(rgb)
    loss = lib().orc_charbonnier(C.c_int64(rgb.shape[0]), _p(rgb), _p(target), C.c_float(eps), _p(g))
    return float(loss), g


def s3im(src, tar, index, patch_h=32, ksize=4, stride=4, mult=1.0, want_grad=True):
    """-> (mult * loss, mult * dloss/dsrc [R,3])"""
    src, tar = _f32(src), _f32(tar)
    index = np.ascontiguousarray(index, np.int64)
    g = np.zeros_like(src) if want_grad else None
    lib().orc_s3im.restype = C.c_double
    loss = lib().orc_s3im(C.c_int64(src.shape[0]), C.c_int64(index.size), _p(index), _p(src), _p(tar),
                          C.c_int(patch_h), C.c_int(ksize), C.c_int(stride), C.c_double(mult), _p(g))
    return float(loss), g


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step):
    """In place on param / exp_avg / exp_avg_sq (float32, contiguous)."""
    lib().orc_adam_step(C.c_int64(param.size), _p(param), _p(_f32(grad)), _p(exp_avg), _p(exp_avg_sq),
                        C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps), C.c_int64(step))


# ---------------------------------------------------------------- ray generation
def generate_rays(cam_idx, coords_yx, c2w, fx, fy, cx, cy):
    """Cameras.generate_rays, perspective / no distortion (nerfstudio/cameras/cameras.py:583-727)."""
    cam_idx = np.ascontiguousarray(cam_idx, np.int64).reshape(-1)
    n = cam_idx.shape[0]
    coords_yx, c2w = _f32(coords_yx).reshape(n, 2), _f32(c2w)
    fx, fy, cx, cy = (_f32(a).reshape(-1) for a in (fx, fy, cx, cy))
    out = {k: np.zeros((n, 3), np.float32) for k in ("origins", "directions", "lookat")}
    out["pixel_area"] = np.zeros(n, np.float32)
    out["dir_norm"] = np.zeros(n, np.float32)
    lib().orc_generate_rays(C.c_int64(n), _p(cam_idx), _p(coords_yx), _p(c2w), _p(fx), _p(fy), _p(cx), _p(cy),
                            _p(out["origins"]), _p(out["directions"]), _p(out["lookat"]), _p(out["pixel_area"]),
                            _p(out["dir_norm"]))
    return out


# ---------------------------------------------------------------- cold sampler queries
def points_anchors(rays_o, rays_d, t_cur, tree_nodes_blob):
    rays_o, rays_d, t_cur = _f32(rays_o), _f32(rays_d), _f32(t_cur)
    R, S = t_cur.shape[0], t_cur.shape[1]
    nodes = np.ascontiguousarray(tree_nodes_blob, np.uint8)
    out = np.zeros((R, S), np.int64)
    lib().orc_points_anchors(C.c_int64(R), C.c_int64(S), _p(rays_o), _p(rays_d), _p(t_cur), _p(nodes),
                             C.c_int64(nodes.size // 128), _p(out))
    return out


def edge_samples(edge_pool_blob, pers_trans_blob, edge_idx, edge_coords):
    edge_idx = np.ascontiguousarray(edge_idx, np.int64)
    edge_coords = _f32(edge_coords)
    n = edge_idx.shape[0]
    pts, idx = np.zeros((n, 2, 3), np.float32), np.zeros((n, 2), np.int64)
    lib().orc_edge_samples(C.c_int64(n), _p(np.ascontiguousarray(edge_pool_blob, np.uint8)),
                           _p(np.ascontiguousarray(pers_trans_blob, np.uint8)), _p(edge_idx), _p(edge_coords), _p(pts),
                           _p(idx))
    return pts, idx


def mark_invisible_nodes(tree_nodes_blob, intri, w2c, bounds):
    """in place on a copy: the node blob with trans_idx = -1 for every node no camera sees (PersSampler_cuda.cu:680-742)"""
    nodes = np.ascontiguousarray(tree_nodes_blob, np.uint8).copy()
    intri, w2c, bounds = _f32(intri), _f32(w2c), _f32(bounds)
    lib().orc_mark_invisible_nodes(C.c_int64(nodes.size // 128), C.c_int64(w2c.shape[0]), _p(nodes), _p(intri),
                                   _p(w2c), _p(bounds))
    return nodes


def set_block_idxs(tree_nodes_blob, centers):
    """a copy of the node blob with block_idx = nearest block centre (SetBlockIdxsNearestKernel, :746-766)"""
    nodes = np.ascontiguousarray(tree_nodes_blob, np.uint8).copy()
    centers = _f32(centers)
    lib().orc_set_block_idxs(C.c_int64(nodes.size // 128), C.c_int64(centers.shape[0]), _p(nodes), _p(centers))
    return nodes


def error_map_update(error_map, indices, pred_rgb, gt_rgb):
    """gfnerf/gf_pipeline.py:180-185 + nerfstudio/data/utils/dataloaders.py:140-142 in numpy: error = sum_c |gt - pred|
    (fp32, channel order 0,1,2), error_map[idx0, idx1, idx2] = error; in place on error_map [n,h,w]; returns error."""
    d = np.abs(np.asarray(gt_rgb, np.float32) - np.asarray(pred_rgb, np.float32))
    err = ((d[:, 0] + d[:, 1]).astype(np.float32) + d[:, 2]).astype(np.float32)
    idx = np.asarray(indices, np.int64)
    error_map[idx[:, 0], idx[:, 1], idx[:, 2]] = err
    return err
