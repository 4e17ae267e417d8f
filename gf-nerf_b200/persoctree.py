"""PersOctree: host-side construction and maintenance of the perspective-warped octree.

Mirrors the C++ class `PersOctree` of the reference (gfnerf/bindings/PtsSampler/PersSampler.cpp:
GetVisiCams :45-88, DistanceSummary :12-26, ctor :92-152, ProcOctree :154-417,
ConstructTreeNode :516-591, PCA :593-610, ConstructTrans :613-831) and produces the same state
blobs (128-byte TreeNode, 576-byte TransInfo; PersSampler.h:31-49), so checkpoints stay
interchangeable with the reference.  The reference's unused Python restatement
gfnerf/persoctree.py has the same role.

This is init-time / every-1000-steps host work (the reference runs it on the CPU too, with a
full D2H/H2D of the node blob); the per-step work -- traversal and marching over these blobs --
is CUDA (csrc/sampler.cu).  numpy only; nothing here touches the GPU.
"""
from typing import List, Tuple

import numpy as np

N_PROS = 12
INIT_NODE_STAT = 1000

TREE_NODE_DTYPE = np.dtype({
    "names": ["center", "side_len", "parent", "childs", "is_leaf_node", "trans_idx", "block_idx"],
    "formats": [("<f4", 3), "<f4", "<i8", ("<i8", 8), "u1", "<i8", "<i8"],
    "offsets": [0, 12, 16, 24, 88, 96, 104],
    "itemsize": 128,
})
TRANS_INFO_DTYPE = np.dtype({
    "names": ["w2xz", "weight", "center", "side_len", "dis_summary"],
    "formats": [("<f4", (N_PROS, 2, 4)), ("<f4", (3, N_PROS)), ("<f4", 3), "<f4", "<f4"],
    "offsets": [0, 384, 528, 540, 544],
    "itemsize": 576,
})


def _byte_copy(arr: np.ndarray) -> np.ndarray:
    """copy of a padded structured array that keeps the padding bytes (zero) instead of leaving them undefined"""
    return np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy().view(arr.dtype)


def _take(arr: np.ndarray, mask: np.ndarray) -> np.ndarray:
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(arr.shape[0], arr.dtype.itemsize)
    return raw[mask].copy().reshape(-1).view(arr.dtype)


def search_order_table() -> np.ndarray:
    """uint8[64]: children in front-to-back order for each ray octant (PersSampler.cpp:137-151).

    The reference sorts 0..7 with cmp(a,b) = (a&bt)^(st&bt), bt = lowest differing bit: a comes first when
    its bit differs from the octant's, i.e. descending in bitrev3(a ^ st)."""
    def bitrev3(v):
        return ((v & 1) << 2) | (v & 2) | ((v >> 2) & 1)
    out = np.zeros(64, np.uint8)
    for st in range(8):
        out[st * 8:(st + 1) * 8] = sorted(range(8), key=lambda a: -bitrev3(a ^ st))
    return out


def distance_summary(dis: np.ndarray) -> float:
    """DistanceSummary (:12-26): geometric mean of the distances below the 25 % quantile (in log space)."""
    dis = np.asarray(dis, np.float32).reshape(-1)
    if dis.size == 0:
        return 1e8
    log_dis = np.log(dis)
    thres = np.float32(np.quantile(log_dis.astype(np.float64), 0.25))
    mask = log_dis < thres
    if mask.sum() < 1:
        return float(np.exp(log_dis.mean()))
    return float(np.exp(log_dis[mask].mean()))


class PersOctree:
    """State: `nodes` (structured TREE_NODE_DTYPE), `trans` (TRANS_INFO_DTYPE), node statistics."""

    def __init__(self, max_depth: int, bbox_side_len: float, split_dist_thres: float, c2w: np.ndarray,
                 intri: np.ndarray, bound: np.ndarray, seed: int = 0, n_rand_pts: int = 32 * 32 * 32,
                 visi_res_w: int = 128):
        self.max_depth = int(max_depth)
        self.bbox_side_len = float(bbox_side_len)
        self.split_dist_thres = float(split_dist_thres)
        self.c2w = np.ascontiguousarray(c2w, np.float32)          # [n,3,4]
        self.intri = np.ascontiguousarray(intri, np.float32)      # [n,3,3]
        self.bound = np.ascontiguousarray(bound, np.float32)      # [n,2]
        self.rng = np.random.RandomState(seed)
        self.n_rand_pts = int(n_rand_pts)
        self._nodes: List[dict] = []
        self._trans: List[np.ndarray] = []
        self._setup_visibility_rays(visi_res_w)
        self._nodes.append(self._blank_node(parent=-1))
        self._construct_tree_node(0, 0, np.zeros(3, np.float32), np.float32(bbox_side_len),
                                  np.arange(self.c2w.shape[0]))
        self.nodes = self._pack_nodes(self._nodes)
        self.trans = np.array(self._trans, dtype=TRANS_INFO_DTYPE).reshape(-1)
        del self._nodes, self._trans
        n = self.nodes.shape[0]
        self.weight_stats = np.full(n, INIT_NODE_STAT, np.int64)
        self.alpha_stats = np.full(n, INIT_NODE_STAT, np.int64)
        self.visit_cnt = np.zeros(n, np.int64)
        self.search_order = search_order_table()

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _blank_node(parent):
        return dict(center=np.zeros(3, np.float32), side_len=np.float32(0), parent=parent, childs=[-1] * 8,
                    is_leaf_node=False, trans_idx=-1, block_idx=-1)

    @staticmethod
    def _pack_nodes(nodes: List[dict]) -> np.ndarray:
        out = np.zeros(len(nodes), TREE_NODE_DTYPE)
        for i, nd in enumerate(nodes):
            out[i]["center"] = nd["center"]
            out[i]["side_len"] = nd["side_len"]
            out[i]["parent"] = nd["parent"]
            out[i]["childs"] = nd["childs"]
            out[i]["is_leaf_node"] = nd["is_leaf_node"]
            out[i]["trans_idx"] = nd["trans_idx"]
            out[i]["block_idx"] = nd["block_idx"]
        return out

    def _setup_visibility_rays(self, res_w):
        """Per-camera 128 x ~72 pixel grid of world-space ray directions (GetVisiCams :51-66)."""
        it = self.intri[0]
        half_w, half_h = float(it[0, 2]), float(it[1, 2])
        cx, cy, fx, fy = float(it[0, 2]), float(it[1, 2]), float(it[0, 0]), float(it[1, 1])
        res_h = int(round(res_w / half_w * half_h))
        i = np.linspace(.5, half_h * 2. - .5, res_h, dtype=np.float32)
        j = np.linspace(.5, half_w * 2. - .5, res_w, dtype=np.float32)
        ii, jj = np.meshgrid(i, j, indexing="ij")
        ii, jj = ii.reshape(-1), jj.reshape(-1)
        cam = np.stack([(jj - cx) / fx, -(ii - cy) / fy, -np.ones_like(jj)], -1).astype(np.float32)   # [n_pix,3]
        self._rays_d = np.einsum("nij,pj->npi", self.c2w[:, :3, :3], cam).astype(np.float32)        # [n_cams,n_pix,3]
        self._cam_pos = self.c2w[:, :3, 3].copy()
        sel = np.zeros((res_h, res_w), bool)
        sel[::4, ::4] = True
        self._coarse_pix = np.nonzero(sel.reshape(-1))[0]
        corner = np.array([half_w / fx, half_h / fy, 1.0])
        self._half_diag_fov = float(np.arccos(1.0 / np.linalg.norm(corner))) + 1e-3

    def _hits(self, cams, pix, lo, hi):
        """cameras of `cams` with at least one ray of the pixel subset `pix` through the box [lo, hi]"""
        d = self._rays_d[cams][:, pix] if pix is not None else self._rays_d[cams]      # [c,p,3]
        o = self._cam_pos[cams][:, None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            a = (lo[None, None] - o) / d
            b = (hi[None, None] - o) / d
        a = np.nan_to_num(a, nan=0., posinf=1e6, neginf=-1e6)
        b = np.nan_to_num(b, nan=0., posinf=1e6, neginf=-1e6)
        far = np.maximum(a, b).min(-1)
        near = np.minimum(a, b).max(-1)
        far = np.minimum(far, self.bound[cams, 1][:, None])
        near = np.maximum(near, self.bound[cams, 0][:, None])
        return (far > near).any(-1)

    def _visible_cams(self, side_len, center, candidates):
        """GetVisiCams (:45-88): cameras with at least one of their 128 x ~72 grid rays through the box inside
        [near, far].  Same predicate, evaluated lazily: only `candidates` (a child box lies inside its
        parent's), a conservative view-cone reject, then a 1/16 subset of the rays before the full grid."""
        if candidates.size == 0:
            return candidates
        lo, hi = (center - side_len * .5).astype(np.float32), (center + side_len * .5).astype(np.float32)
        rel = center[None] - self._cam_pos[candidates]
        dist = np.linalg.norm(rel, axis=-1)
        radius = side_len * 0.8660254 + 1e-6
        fwd = -self.c2w[candidates][:, :3, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            cosang = np.clip((rel * fwd).sum(-1) / dist, -1, 1)
            ang = np.arccos(cosang) - np.arcsin(np.clip(radius / dist, 0, 1))
        maybe = (dist <= radius) | (ang <= self._half_diag_fov)
        cams = candidates[maybe]
        if cams.size == 0:
            return cams
        vis = self._hits(cams, self._coarse_pix, lo, hi)
        rest = np.nonzero(~vis)[0]
        for s0 in range(0, rest.size, 64):
            sub = rest[s0:s0 + 64]
            vis[sub] = self._hits(cams[sub], None, lo, hi)
        return cams[vis]

    # ------------------------------------------------------------------ build
    def _construct_tree_node(self, u, depth, center, side_len, candidates):
        nd = self._nodes[u]
        nd["center"] = center.astype(np.float32)
        nd["side_len"] = np.float32(side_len)
        if depth > self.max_depth:
            nd["is_leaf_node"] = True
            return
        visi = self._visible_cams(np.float32(side_len), center, candidates)
        cam_dis = np.linalg.norm(self._cam_pos[visi] - center[None], axis=-1).astype(np.float32)
        dsum = distance_summary(cam_dis)
        unaddressed = visi.size >= N_PROS // 2 and dsum < side_len * self.split_dist_thres
        if unaddressed:
            for st in range(8):
                v = len(self._nodes)
                self._nodes.append(self._blank_node(parent=u))
                off = np.array([((st >> 2) & 1) - .5, ((st >> 1) & 1) - .5, (st & 1) - .5], np.float32)
                nd["childs"][st] = v
                self._construct_tree_node(v, depth + 1, (center + side_len * np.float32(.5) * off).astype(np.float32),
                                          np.float32(side_len * .5), visi)
        elif visi.size < N_PROS // 2:
            nd["is_leaf_node"] = True          # leaf, but invalid: not enough visible cameras
        else:
            nd["is_leaf_node"] = True
            nd["trans_idx"] = len(self._trans)
            rand_pts = ((self.rng.rand(self.n_rand_pts, 3).astype(np.float32) - .5) * side_len + center[None]).astype(np.float32)
            tr = self.construct_trans(rand_pts, self.c2w[visi], self.intri[0], center)
            tr["side_len"] = side_len
            self._trans.append(tr)

    def construct_trans(self, rand_pts, c2w, intri, center) -> np.ndarray:
        """ConstructTrans (:613-831): 6 well-spread cameras re-aimed at the cell centre -> 12 (x/z, y/z)
        projections; PCA of the projected sample points -> 3 x 12 mixing weights, normalised by the mean
        inverse Jacobian so that one warp-space unit is about one pixel-ish step."""
        n_virt = N_PROS // 2
        n_cur = c2w.shape[0]
        cam_pos = c2w[:, :3, 3].astype(np.float32)
        cam_axes = np.linalg.inv(c2w[:, :3, :3].astype(np.float64)).astype(np.float32)
        dis = np.linalg.norm(cam_pos - center[None], axis=-1).astype(np.float32)
        dis_summary = np.float32(distance_summary(dis))
        normed = (cam_pos - center[None]) / dis[:, None]
        dis_pairs = np.linalg.norm(normed[None] - normed[:, None], axis=-1)
        good = [int(self.rng.randint(n_cur))]
        marks = np.zeros(n_cur, bool)
        marks[good[0]] = True
        for _ in range(1, min(n_virt, n_cur)):      # farthest-point selection (:652-673)
            cur = np.where(marks[None, :], dis_pairs, np.float32(1e8)).min(-1)
            cur[marks] = -2.
            candi = int(np.argmax(cur))             # first maximum, like the reference's strict '>'
            marks[candi] = True
            good.append(candi)
        i = 0
        while len(good) < n_virt:
            good.append(good[i])
            i += 1
        good = np.array(good)
        cam_scale = np.clip(dis / dis_summary, 1., 1e9).astype(np.float32)
        rel = (cam_pos - center[None]) / dis[:, None] * np.clip(dis[:, None], dis_summary, 1e9)
        good_rel = rel[good].astype(np.float32)
        good_pos = (good_rel + center[None]).astype(np.float32)
        good_axis = cam_axes[good].copy()
        good_scale = cam_scale[good]
        expect_z = good_rel / np.linalg.norm(good_rel, axis=-1, keepdims=True)
        rots = np.zeros((n_virt, 3, 3), np.float32)
        for k in range(n_virt):                    # rotate each camera so that its z axis looks along expect_z
            fz, tz = good_axis[k, 2], expect_z[k]
            crossed = np.cross(fz, tz)
            cos_v = float(np.clip(np.dot(fz, tz), -0.999999, 0.999999))
            sin_v = float(np.clip(np.linalg.norm(crossed), -0.999999, 0.999999))
            angle = np.arcsin(sin_v)
            if cos_v < 0:
                angle = np.pi - angle
            nrm = np.linalg.norm(crossed)
            axis = crossed / nrm if nrm > 0 else crossed
            K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]], np.float64)
            rots[k] = (np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)).astype(np.float32)  # AngleAxis
        good_axis = np.matmul(good_axis, rots.transpose(0, 2, 1))
        focal = np.float32(intri[0, 0] / intri[0, 2])
        x_axis = good_axis[:, 0] * focal * good_scale[:, None]
        y_axis = good_axis[:, 1] * focal * good_scale[:, None]
        z_axis = good_axis[:, 2]
        xa = np.concatenate([x_axis, y_axis], 0)
        za = np.concatenate([z_axis, z_axis], 0)
        wp = np.concatenate([good_pos, good_pos], 0)
        frame = np.zeros((N_PROS, 2, 4), np.float32)
        frame[:, 0, :3] = xa
        frame[:, 1, :3] = za
        frame[:, 0, 3] = -(xa * wp).sum(-1)
        frame[:, 1, 3] = -(za * wp).sum(-1)
        # projected sample points and their Jacobians (:779-812)
        tp = np.einsum("pij,nj->npi", frame[:, :, :3], rand_pts) + frame[None, :, :, 3]     # [n,12,2]
        dv_da = 1. / tp[:, :, 1]
        dv_db = tp[:, :, 0] / -np.square(tp[:, :, 1])
        dv_dxyz = dv_da[:, :, None] * frame[None, :, 0, :3] + dv_db[:, :, None] * frame[None, :, 1, :3]   # [n,12,3]
        if not tp[:, :, 1].max() < 0:
            raise RuntimeError("ConstructTrans: sample points behind a virtual camera (PersSampler.cpp:795)")
        v = (tp[:, :, 0] / tp[:, :, 1]).astype(np.float64)                                   # [n,12]
        moved = v - v.mean(0, keepdims=True)
        cov = moved.T @ moved / moved.shape[0]
        L, V = np.linalg.eigh(cov)
        order = np.argsort(-L, kind="stable")
        V = V[:, order].T[:3].astype(np.float32)                                             # [3,12]
        jac = np.einsum("rk,nkc->nrc", V, dv_dxyz)                                           # [n,3,3]
        jac_w2i = np.matmul(dv_dxyz, np.linalg.inv(jac.astype(np.float64)).astype(np.float32))   # [n,12,3]
        mean_step = (1. / np.abs(jac_w2i).max(1)).mean(0)                                    # [3]
        V = (V / mean_step[:, None]).astype(np.float32)
        if not (np.isfinite(V).all() and np.isfinite(frame).all()):
            raise RuntimeError("ConstructTrans: non-finite transform")
        out = np.zeros((), TRANS_INFO_DTYPE)
        out["w2xz"] = frame
        out["weight"] = V
        out["center"] = center
        out["dis_summary"] = dis_summary
        return out

    # ------------------------------------------------------------------ blobs
    def tree_nodes_blob(self) -> np.ndarray:
        return np.ascontiguousarray(self.nodes).view(np.uint8).reshape(-1).copy()

    def pers_trans_blob(self) -> np.ndarray:
        return np.ascontiguousarray(self.trans).view(np.uint8).reshape(-1).copy()

    def load_blobs(self, tree_nodes: np.ndarray, pers_trans: np.ndarray):
        # byte copies: a field-wise copy of a padded structured array would leave the padding undefined
        self.nodes = np.ascontiguousarray(tree_nodes, np.uint8).reshape(-1).copy().view(TREE_NODE_DTYPE)
        self.trans = np.ascontiguousarray(pers_trans, np.uint8).reshape(-1).copy().view(TRANS_INFO_DTYPE)

    # ------------------------------------------------------------------ maintenance
    def proc_octree(self, compact: bool, subdivide: bool, brute_force: bool) -> None:
        """ProcOctree (:154-417): prune invalid leaves, collapse childless / single-child interior nodes,
        renumber (order of first appearance), then split every visited leaf into 8 (depth-first order).
        `nodes`, `weight_stats`, `alpha_stats`, `visit_cnt` must hold the device's current values."""
        nb = _byte_copy(self.nodes)
        n_before = nb.shape[0]
        w_before, a_before, visit = self.weight_stats, self.alpha_stats, self.visit_cnt
        childs, parent, leaf, tidx = nb["childs"], nb["parent"], nb["is_leaf_node"], nb["trans_idx"]
        while compact:
            for u in range(n_before):
                if not leaf[u]:
                    continue
                if tidx[u] < 0 and parent[u] >= 0:
                    v = parent[u]
                    childs[v][childs[v] == u] = -1
            update = False
            for u in range(1, n_before):
                if not (childs[u] >= 0).any():
                    if not leaf[u]:
                        update = True
                    leaf[u] = 1
            if not update:
                break
        if compact:
            def single_child(x):
                idx = np.nonzero(childs[x] >= 0)[0]
                return int(idx[0]) if idx.size == 1 else -1
            for u in range(n_before):
                if leaf[u] and tidx[u] < 0:
                    continue
                v = parent[u]
                while v >= 0 and parent[v] >= 0 and single_child(v) >= 0:
                    vv = parent[v]
                    childs[vv][childs[vv] == v] = u
                    parent[u] = vv
                    tidx[v] = -1
                    leaf[v] = 1
                    v = vv
        keep = (~leaf.astype(bool)) | (tidx >= 0)
        new_idx = np.full(n_before, -1, np.int64)
        new_idx[keep] = np.arange(int(keep.sum()))
        inv_idx = np.nonzero(keep)[0]
        assert new_idx[0] == 0
        nn = _take(nb, keep)
        pm = nn["parent"] >= 0
        nn["parent"][pm] = new_idx[nn["parent"][pm]]
        cm = nn["childs"] >= 0
        nn["childs"][cm] = new_idx[nn["childs"][cm]]
        if (nn["parent"][pm] < 0).any() or (nn["childs"][cm] < 0).any():
            # the reference's CHECK_GE(node.parent, 0) / CHECK_GE(node.childs[st], 0), PersSampler.cpp:307,315: a
            # pruned leaf still linked from its parent, i.e. ProcOctree without `compact` on a tree with pruned leaves
            raise RuntimeError("ProcOctree: a removed node is still linked (call it with compact=True)")
        nw, na = w_before[keep].copy(), a_before[keep].copy()
        if subdivide:
            out_nodes, out_w, out_a = [], [], []
            src = nn

            def rec(u, pa):
                new_u = len(out_nodes)
                node = np.zeros((), TREE_NODE_DTYPE)
                for f in TREE_NODE_DTYPE.names:
                    node[f] = src[u][f]
                node["parent"] = pa
                out_nodes.append(node)
                out_w.append(int(nw[u]))
                out_a.append(int(na[u]))
                if node["is_leaf_node"]:
                    if not brute_force and visit[inv_idx[u]] <= 4:
                        return new_u
                    for st in range(8):
                        off = np.array([((st >> 2) & 1) - .5, ((st >> 1) & 1) - .5, (st & 1) - .5], np.float32)
                        v = len(out_nodes)
                        ch = np.zeros((), TREE_NODE_DTYPE)
                        ch["center"] = node["center"] + node["side_len"] * np.float32(.5) * off
                        ch["side_len"] = node["side_len"] * np.float32(.5)
                        ch["parent"] = new_u
                        ch["childs"] = -1
                        ch["is_leaf_node"] = 1
                        ch["trans_idx"] = node["trans_idx"]
                        # block_idx is left unset by the reference (:376); 0 here
                        out_nodes.append(ch)
                        out_w.append(out_w[new_u])
                        out_a.append(out_a[new_u])
                        out_nodes[new_u]["childs"][st] = v
                    out_nodes[new_u]["is_leaf_node"] = 0
                    out_nodes[new_u]["trans_idx"] = -1
                    out_w[new_u] = INIT_NODE_STAT
                    out_a[new_u] = INIT_NODE_STAT
                else:
                    for st in range(8):
                        c = int(out_nodes[new_u]["childs"][st])
                        if c >= 0:
                            out_nodes[new_u]["childs"][st] = rec(c, new_u)
                return new_u

            import sys
            lim = sys.getrecursionlimit()
            sys.setrecursionlimit(max(lim, 10000))
            try:
                rec(0, -1)
            finally:
                sys.setrecursionlimit(lim)
            nn = np.zeros(len(out_nodes), TREE_NODE_DTYPE)
            for i_, nd_ in enumerate(out_nodes):
                for f in TREE_NODE_DTYPE.names:
                    nn[i_][f] = nd_[f]
            nw, na = np.array(out_w, np.int64), np.array(out_a, np.int64)
        self.nodes = nn
        self.weight_stats, self.alpha_stats = nw, na
        self.visit_cnt = np.zeros(nn.shape[0], np.int64)

    # ------------------------------------------------------------------ stats
    def n_valid_leaves(self) -> int:
        return int((self.nodes["trans_idx"] >= 0).sum())


def aerial_rig(n_side: int = 20, height: float = 2.0, extent: float = 4.0, jitter_deg: float = 15.0,
               fx: float = 1000.0, width: int = 1920, height_px: int = 1080, near: float = 0.01, far: float = 512.0,
               seed: int = 0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Synthetic aerial camera rig of SURVEY.md 8(d): n_side^2 cameras on a grid at z=`height` over a
    [-extent, extent]^2 ground patch, looking down with up to `jitter_deg` of tilt.  OpenGL camera convention
    (looks along -z, like nerfstudio).  Returns c2w [n,3,4], intri [n,3,3], bounds [n,2]."""
    rng = np.random.RandomState(seed)
    xs = np.linspace(-extent, extent, n_side)
    c2w = []
    for x in xs:
        for y in xs:
            tilt = np.deg2rad(rng.uniform(-jitter_deg, jitter_deg, size=2))
            look = np.array([np.tan(tilt[0]), np.tan(tilt[1]), -1.0])
            look /= np.linalg.norm(look)
            zc = -look                                # camera +z points away from the view direction
            xc = np.cross(np.array([0., 1., 0.]), zc)
            xc /= np.linalg.norm(xc)
            yc = np.cross(zc, xc)
            m = np.zeros((3, 4))
            m[:, 0], m[:, 1], m[:, 2], m[:, 3] = xc, yc, zc, np.array([x, y, height])
            c2w.append(m)
    c2w = np.stack(c2w).astype(np.float32)
    n = c2w.shape[0]
    intri = np.tile(np.array([[fx, 0, width / 2], [0, fx, height_px / 2], [0, 0, 1]], np.float32)[None], (n, 1, 1))
    bounds = np.tile(np.array([[near, far]], np.float32), (n, 1))
    return c2w, intri, bounds


def rig_rays(c2w: np.ndarray, intri: np.ndarray, n_rays: int, seed: int = 0):
    """Random (camera, pixel) rays of a rig: origins [R,3], unit directions [R,3], camera index [R]."""
    rng = np.random.RandomState(seed)
    cam = rng.randint(0, c2w.shape[0], size=n_rays)
    w, h = intri[0, 0, 2] * 2, intri[0, 1, 2] * 2
    px = rng.uniform(0, w, size=n_rays).astype(np.float32)
    py = rng.uniform(0, h, size=n_rays).astype(np.float32)
    fx, fy, cx, cy = intri[cam, 0, 0], intri[cam, 1, 1], intri[cam, 0, 2], intri[cam, 1, 2]
    d_cam = np.stack([(px - cx) / fx, -(py - cy) / fy, -np.ones_like(px)], -1).astype(np.float32)
    d = np.einsum("nij,nj->ni", c2w[cam][:, :, :3], d_cam)
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    o = c2w[cam][:, :, 3].astype(np.float32)
    return np.ascontiguousarray(o), np.ascontiguousarray(d), cam.astype(np.int64)


def frame_rays(c2w_cam: np.ndarray, intri_cam: np.ndarray, width: int, height: int):
    """All pixel-centre rays of one camera, row-major [H*W]: origins, unit directions (same pinhole / OpenGL
    convention as rig_rays; the caller of the reference is Cameras.generate_rays, out of scope)."""
    px, py = np.meshgrid(np.arange(width, dtype=np.float32) + 0.5, np.arange(height, dtype=np.float32) + 0.5)
    fx, fy, cx, cy = intri_cam[0, 0], intri_cam[1, 1], intri_cam[0, 2], intri_cam[1, 2]
    d_cam = np.stack([(px - cx) / fx, -(py - cy) / fy, -np.ones_like(px)], -1).reshape(-1, 3).astype(np.float32)
    d = d_cam @ c2w_cam[:, :3].T.astype(np.float32)
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    o = np.broadcast_to(c2w_cam[:, 3].astype(np.float32), d.shape)
    return np.ascontiguousarray(o), np.ascontiguousarray(d)
