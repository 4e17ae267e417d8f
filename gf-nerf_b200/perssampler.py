"""PersSampler: host-side mirror of the reference sampler operator.

`PersSamplerCore` mirrors the C++ class behind `torch.classes.my_classes.PersSampler`
(reference gfnerf/bindings/PtsSampler/PersSampler.cpp:899-1016, PersSampler_cuda.cu:321-477,
584-677, bindings.cpp:18-299): same constructor arguments (`InitSampler`), method names, return
orders and state blobs.  `PersSampler` mirrors the nn.Module shell (reference
gfnerf/perssampler.py:47-447).  Traversal / marching / voting run in the C-ABI library
(csrc/sampler.cu); octree construction and `ProcOctree` are host work (persoctree.py), as in the
reference.  There is no CPU fallback for the per-step path.

Besides the reference's dense `[R,1024,...]` API there is `sample_compact`, the CSR layout the
fused engine uses: only valid samples exist, 36 B each instead of 76 B x 1024 slots per ray.
"""
import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from .persoctree import INIT_NODE_STAT, PersOctree

MAX_SAMPLE_PER_RAY = 1024


@dataclass
class CompactSamples:
    """CSR sample set of one ray batch (device tensors, capacity R*1024; `total` valid entries)."""
    n_rays: int
    counts: torch.Tensor        # int32 [R]
    offsets: torch.Tensor       # int32 [R+1]
    total: torch.Tensor         # int32 [1] (device)
    pts01: torch.Tensor         # f32 [cap,3]   (warp+1.5)/3
    anchor: torch.Tensor        # int32 [cap]   trans_idx
    node: torch.Tensor          # int32 [cap]   tree node
    t: torch.Tensor             # f32 [cap]
    delta: torch.Tensor         # f32 [cap]
    ray_id: torch.Tensor        # int32 [cap]
    first_oct_dis: torch.Tensor  # f32 [R]
    rays_d_unit: torch.Tensor   # f32 [R,3]  normalised (what the march uses, PersSampler_cuda.cu:323)
    rays_d: torch.Tensor        # f32 [R,3]  the bundle's directions AS GIVEN: what the field's SH encodes -- the reference
                                # normalises only inside GetSamples; frustums.directions stay the bundle's
                                # (perssampler.py:414-418, nerfacto_field.py:518-521; pinned by tests/golden/ref_model.npz)


def build_octree(max_depth: int, bbox_side_len: float, split_dist_thres: float, c2w, intri, bounds, seed: int = 0,
                 n_rand_pts: int = 32 * 32 * 32, visi_res_w: int = 128) -> PersOctree:
    """PersOctree::PersOctree (PersSampler.cpp:92-152) through the C++ host builder gf_octree_build
    (csrc/octree_build.cu); returns the blobs wrapped in a `PersOctree` (whose numpy constructor is the independent
    restatement the builder is tested against)."""
    import ctypes as C
    c2w = np.ascontiguousarray(c2w, np.float32)
    intri = np.ascontiguousarray(intri, np.float32)
    bounds = np.ascontiguousarray(bounds, np.float32)
    if c2w.ndim != 3 or c2w.shape[1:] != (3, 4) or intri.shape != (c2w.shape[0], 3, 3) or bounds.shape != (c2w.shape[0], 2):
        raise RuntimeError("build_octree: c2w f32 [n,3,4], intri f32 [n,3,3], bounds f32 [n,2]")
    L = _lib.lib()
    handle, n_nodes, n_trans = C.c_void_p(), C.c_int64(), C.c_int64()
    _lib.check(L.gf_octree_build(int(max_depth), float(bbox_side_len), float(split_dist_thres), c2w.ctypes.data,
                                 intri.ctypes.data, bounds.ctypes.data, c2w.shape[0], int(seed) & 0xffffffff,
                                 int(n_rand_pts), int(visi_res_w), C.byref(handle), C.byref(n_nodes), C.byref(n_trans)),
               "gf_octree_build")
    nodes = np.empty(n_nodes.value * 128, np.uint8)
    trans = np.empty(n_trans.value * 576, np.uint8)
    _lib.check(L.gf_octree_build_fetch(handle, nodes.ctypes.data, trans.ctypes.data), "gf_octree_build_fetch")
    oc = PersOctree.__new__(PersOctree)
    oc.max_depth, oc.bbox_side_len, oc.split_dist_thres = int(max_depth), float(bbox_side_len), float(split_dist_thres)
    oc.c2w, oc.intri, oc.bound = c2w, intri, bounds
    oc.load_blobs(nodes, trans)
    n = oc.nodes.shape[0]
    oc.weight_stats = np.full(n, INIT_NODE_STAT, np.int64)
    oc.alpha_stats = np.full(n, INIT_NODE_STAT, np.int64)
    oc.visit_cnt = np.zeros(n, np.int64)
    so = np.zeros(64, np.uint8)
    _lib.check(L.gf_octree_search_order(so.ctypes.data), "gf_octree_search_order")
    oc.search_order = so
    return oc


class PersSamplerCore:
    def __init__(self):
        self.octree: Optional[PersOctree] = None

    # ---- construction (PersSampler::PersSampler, PersSampler.cpp:899-952) -------------------
    def InitSampler(self, split_dist_thres, sub_div_milestones, compact_freq, max_oct_intersect_per_ray, global_near,
                    scale_by_dis, bbox_levels, sample_l, max_level, c2w, w2c, intri, bounds, mode,
                    sampled_oct_per_ray, ray_march_fineness, ray_march_init_fineness,
                    ray_march_fineness_decay_end_iter, device=None, seed: int = 0, octree: PersOctree = None):
        if not torch.cuda.is_available():
            raise RuntimeError("PersSampler needs a CUDA device: the B200 kernels have no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.sub_div_milestones_ = [int(v) for v in sub_div_milestones][::-1]      # reversed, popped from the back
        self.compact_freq_ = int(compact_freq)
        self.max_oct_intersect_per_ray_ = int(max_oct_intersect_per_ray)
        self.global_near_ = float(np.float32(global_near))
        self.scale_by_dis_ = bool(scale_by_dis)
        self.sample_l_ = float(np.float32(sample_l))
        self.mode_ = int(mode)
        self.sampled_oct_per_ray_ = float(sampled_oct_per_ray)
        self.ray_march_fineness_ = float(ray_march_fineness)
        self.ray_march_init_fineness_ = float(ray_march_init_fineness)
        self.ray_march_fineness_decay_end_iter_ = float(ray_march_fineness_decay_end_iter)
        to_np = lambda x: x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
        self.w2c_ = torch.as_tensor(to_np(w2c), dtype=torch.float32).to(self.device).contiguous()
        self.intri_ = torch.as_tensor(to_np(intri), dtype=torch.float32).to(self.device).contiguous()
        self.bound_ = torch.as_tensor(to_np(bounds), dtype=torch.float32).to(self.device).contiguous()
        if octree is None:
            octree = build_octree(int(max_level), float(1 << (int(bbox_levels) - 1)), float(np.float32(split_dist_thres)),
                                  to_np(c2w), to_np(intri), to_np(bounds), seed=seed)
        self.octree = octree
        self._edge_pool = None
        self.n_volumes_ = int(octree.trans.shape[0])
        self._upload_octree(stats=True)
        self.search_order_ = torch.from_numpy(octree.search_order.copy()).to(self.device)
        self._ws = {}
        self.generator = None   # optional torch.Generator (cuda) for the march noise

    def _upload_octree(self, stats: bool):
        oc = self.octree
        self.tree_nodes_gpu_ = torch.from_numpy(oc.tree_nodes_blob()).to(self.device)
        self.pers_trans_gpu_ = torch.from_numpy(oc.pers_trans_blob()).to(self.device)
        if stats:
            self.tree_weight_stats_ = torch.from_numpy(oc.weight_stats.copy()).to(self.device)
            self.tree_alpha_stats_ = torch.from_numpy(oc.alpha_stats.copy()).to(self.device)
            self.tree_visit_cnt_ = torch.from_numpy(oc.visit_cnt.copy()).to(self.device)

    @property
    def n_nodes(self) -> int:
        return self.tree_nodes_gpu_.numel() // 128

    # ---- workspace ---------------------------------------------------------------------
    def _buf(self, name, shape, dtype):
        t = self._ws.get(name)
        n = int(np.prod(shape))
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(n, dtype=dtype, device=self.device)
            self._ws[name] = t
        return t[:n].view(*shape)

    def _noise(self, n_rays: int) -> torch.Tensor:
        """PersSampler_cuda.cu:380-389: ones in VALIDATE mode, U(0.5,1.5) in TRAIN mode, times the fineness."""
        n = MAX_SAMPLE_PER_RAY + n_rays + 10
        if self.mode_ == 1:
            noise = torch.ones(n, dtype=torch.float32, device=self.device)
        else:
            noise = torch.rand(n, dtype=torch.float32, device=self.device, generator=self.generator) - .5 + 1.
        return noise.mul_(float(np.float32(self.ray_march_fineness_)))

    def _launch(self, rays_o, rays_d_unit, noise, out: "_lib.SamplerOut"):
        import ctypes as C
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gf_sampler_get_samples(
                rays_o.shape[0], _lib.ptr(rays_o), _lib.ptr(rays_d_unit), _lib.ptr(noise),
                _lib.ptr(self.tree_nodes_gpu_), self.n_nodes, _lib.ptr(self.pers_trans_gpu_),
                self.pers_trans_gpu_.numel() // 576, _lib.ptr(self.search_order_), self.global_near_, self.sample_l_,
                int(self.scale_by_dis_), self.max_oct_intersect_per_ray_, C.byref(out), _lib.cur_stream()),
                "gf_sampler_get_samples")

    @staticmethod
    def _normalise(rays_o, rays_d):
        rays_o = rays_o.contiguous().float()
        rays_d = (rays_d / torch.linalg.norm(rays_d, 2, -1, True)).contiguous().float()   # :323
        return rays_o, rays_d

    # ---- reference API: dense [R,1024,...] (PersSampler::GetSamples, :321-477) ---------------
    def GetSamples(self, rays_o_raw: torch.Tensor, rays_d_raw: torch.Tensor, bounds_raw: torch.Tensor = None,
                   noise: torch.Tensor = None) -> List[torch.Tensor]:
        _lib.require_cuda(rays_o_raw, rays_d_raw)
        rays_o, rays_d = self._normalise(rays_o_raw, rays_d_raw)
        R, S, dev = rays_o.shape[0], MAX_SAMPLE_PER_RAY, self.device
        train = self.mode_ != 1
        if noise is None:
            noise = self._noise(R)
        world = torch.zeros((R, S, 3), dtype=torch.float32, device=dev)
        warp = torch.zeros((R, S, 3), dtype=torch.float32, device=dev)
        dirs = torch.zeros((R, S, 3), dtype=torch.float32, device=dev)
        anchors = torch.zeros((R, S, 3), dtype=torch.int64, device=dev)
        dists = torch.zeros((R, S), dtype=torch.float32, device=dev)
        ts = torch.zeros((R, S), dtype=torch.float32, device=dev)
        start_end = torch.zeros((R, 2), dtype=torch.int64, device=dev)
        first = torch.zeros((R, 1), dtype=torch.float32, device=dev)
        counts = torch.zeros(R, dtype=torch.int32, device=dev)
        n_oct = torch.zeros(R, dtype=torch.int32, device=dev) if train else None
        out = _lib.SamplerOut(world_pts=_lib.ptr(world), warp_pts=_lib.ptr(warp), dirs=_lib.ptr(dirs),
                              dists=_lib.ptr(dists), ts=_lib.ptr(ts), anchors_i64=_lib.ptr(anchors),
                              anchors_i32=None, pts_idx_start_end=_lib.ptr(start_end), counts=_lib.ptr(counts),
                              first_oct_dis=_lib.ptr(first), n_oct=_lib.ptr(n_oct), packed=None)
        if R > 0:
            self._launch(rays_o, rays_d, noise.contiguous(), out)
        self.last_counts_ = counts
        if train and R > 0:   # EMA of leaves per ray (:386-387); the only host read, off the critical path
            self._pending_n_oct = (n_oct, R)
        return [world, warp, dirs, dists, ts, anchors, start_end, first]

    def flush_stats(self):
        """Folds the last batch's leaves-per-ray into sampled_oct_per_ray_ (host sync; call when convenient)."""
        p = getattr(self, "_pending_n_oct", None)
        if p is not None:
            n_oct, R = p
            per_ray = float(n_oct.sum().item()) / float(R)
            self.sampled_oct_per_ray_ = self.sampled_oct_per_ray_ * .9 + per_ray * .1
            self._pending_n_oct = None

    # ---- compact API (fused engine) -------------------------------------------------------
    def sample_compact(self, rays_o_raw, rays_d_raw, noise: torch.Tensor = None, want_n_oct=False,
                       slot: int = 0) -> CompactSamples:
        """`slot` selects one of several independent output workspaces, so that the samples of batch k + 1 can be
        produced (on another stream) while batch k's are still being read by its backward pass."""
        _lib.require_cuda(rays_o_raw, rays_d_raw)
        rays_o, rays_d = self._normalise(rays_o_raw, rays_d_raw)
        R, S = rays_o.shape[0], MAX_SAMPLE_PER_RAY
        cap = max(R * S, 1)
        if noise is None:
            noise = self._noise(R)
        sfx = "@%d" % slot if slot else ""
        packed = self._buf("packed", (cap, 8), torch.float32)      # consumed by gf_sampler_compact right below
        counts = self._buf("counts" + sfx, (R,), torch.int32)
        offsets = self._buf("offsets" + sfx, (R + 1,), torch.int32)
        total = self._buf("total" + sfx, (1,), torch.int32)
        first = self._buf("first" + sfx, (R,), torch.float32)
        n_oct = self._buf("n_oct" + sfx, (R,), torch.int32) if want_n_oct else None
        cs = CompactSamples(
            n_rays=R, counts=counts, offsets=offsets, total=total,
            pts01=self._buf("pts01" + sfx, (cap, 3), torch.float32),
            anchor=self._buf("anchor" + sfx, (cap,), torch.int32),
            node=self._buf("node" + sfx, (cap,), torch.int32), t=self._buf("t" + sfx, (cap,), torch.float32),
            delta=self._buf("delta" + sfx, (cap,), torch.float32),
            ray_id=self._buf("ray_id" + sfx, (cap,), torch.int32),
            first_oct_dis=first, rays_d_unit=rays_d, rays_d=rays_d_raw.contiguous().float())
        if R == 0:
            offsets.zero_()
            total.zero_()
            return cs
        out = _lib.SamplerOut(world_pts=None, warp_pts=None, dirs=None, dists=None, ts=None, anchors_i64=None,
                              anchors_i32=None, pts_idx_start_end=None, counts=_lib.ptr(counts),
                              first_oct_dis=_lib.ptr(first), n_oct=_lib.ptr(n_oct), packed=_lib.ptr(packed))
        import contextlib
        hook = getattr(self, "stage_hook", None) or (lambda name: contextlib.nullcontext())
        with hook("sample_rays"):
            self._launch(rays_o, rays_d, noise.contiguous(), out)
        L, st = _lib.lib(), _lib.cur_stream()
        with torch.cuda.device(self.device):
            with hook("scan"):
                _lib.check(L.gf_sampler_scan_counts(R, _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(total), st),
                           "gf_sampler_scan_counts")
            with hook("compact"):
                _lib.check(L.gf_sampler_compact(R, _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(packed),
                                                _lib.ptr(cs.pts01), _lib.ptr(cs.anchor), _lib.ptr(cs.node),
                                                _lib.ptr(cs.t), _lib.ptr(cs.delta), _lib.ptr(cs.ray_id), st),
                           "gf_sampler_compact")
        if want_n_oct:
            self._pending_n_oct = (n_oct, R)
        return cs

    # ---- training feedback (PersSampler::UpdateOctNodes, :584-677) ------------------------
    def _vote(self, n_rays, counts, offsets, c_node, weights, alphas):
        """Vote (MarkVistNodeKernel), [MAX-reduce the votes over the data-parallel ranks], apply (stat update +
        MarkInvalidNodes).  `vote_reduce` is set by the engine when world_size > 1; without it this is exactly
        gf_sampler_update_oct_nodes."""
        scratch = self._buf("vote", (3 * self.n_nodes,), torch.int64)
        L, st = _lib.lib(), _lib.cur_stream()
        with torch.cuda.device(self.device):
            _lib.check(L.gf_sampler_vote(
                n_rays, _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(c_node), _lib.ptr(weights), _lib.ptr(alphas),
                self.n_nodes, _lib.ptr(self.tree_visit_cnt_), _lib.ptr(scratch), st), "gf_sampler_vote")
            reduce = getattr(self, "vote_reduce", None)
            if reduce is not None:
                reduce([scratch, self.tree_visit_cnt_])
            _lib.check(L.gf_sampler_apply_votes(
                _lib.ptr(self.tree_nodes_gpu_), self.n_nodes, _lib.ptr(self.tree_weight_stats_),
                _lib.ptr(self.tree_alpha_stats_), _lib.ptr(scratch), st), "gf_sampler_apply_votes")

    def _milestones(self, iter_step: int):
        while self.sub_div_milestones_ and self.sub_div_milestones_[-1] <= iter_step:
            self.ProcOctree(True, True, self.sub_div_milestones_[-1] <= 0)
            self.MarkInvisibleNodes()
            self.ProcOctree(True, False, False)
            self.sub_div_milestones_.pop()
        if iter_step % self.compact_freq_ == 0:
            self.ProcOctree(True, False, False)

    def UpdateOctNodes(self, sampled_anchors, pts_idx_bounds, sampled_weight, sampled_alpha, iter_step: int):
        """Dense reference signature: anchors i64 [R,1024,3], bounds i64 [R,1024,2], weights/alphas [R,1024,1]."""
        R = sampled_weight.shape[0]
        se = pts_idx_bounds[:, 0, :]
        counts = (se[:, 1] - se[:, 0]).to(torch.int32).contiguous()
        offsets = (torch.arange(R + 1, device=self.device, dtype=torch.int64) * MAX_SAMPLE_PER_RAY).to(torch.int32)
        c_node = sampled_anchors.reshape(R * MAX_SAMPLE_PER_RAY, 3)[:, 1].to(torch.int32).contiguous()
        self._vote(R, counts, offsets, c_node, sampled_weight.reshape(-1).contiguous().float(),
                   sampled_alpha.reshape(-1).contiguous().float())
        self._milestones(int(iter_step))

    def update_oct_nodes_compact(self, cs: CompactSamples, weights, alphas, iter_step: int):
        self._vote(cs.n_rays, cs.counts, cs.offsets, cs.node, weights, alphas)
        self._milestones(int(iter_step))

    def ProcOctree(self, compact: bool, subdivide: bool, brute_force: bool):
        """PersOctree::ProcOctree (PersSampler.cpp:154-417) on the device: one kernel (gf_octree_proc_device,
        csrc/octree_device.cu) rebuilds the node blob and the statistics in HBM -- byte-identical to the reference's
        host rebuild (gf_octree_proc, kept as the checker: `ProcOctreeHost`) without its D2H / H2D round trip of the
        blob.  The only host read is the new node count + error word (12 bytes)."""
        n_in = self.n_nodes
        cap = 9 * n_in if subdivide else n_in
        L, dev = _lib.lib(), self.device
        scratch_bytes = int(L.gf_octree_proc_device_scratch_bytes(n_in))
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
        nodes_o = torch.empty(cap * 128, dtype=torch.uint8, device=dev)
        w_o = torch.empty(cap, dtype=torch.int64, device=dev)
        a_o = torch.empty(cap, dtype=torch.int64, device=dev)
        res = torch.zeros(4, dtype=torch.int32, device=dev)           # [0:2] = n_out (int64), [2] = error word
        with torch.cuda.device(dev):
            _lib.check(L.gf_octree_proc_device(
                _lib.ptr(self.tree_nodes_gpu_), n_in, _lib.ptr(self.tree_weight_stats_), _lib.ptr(self.tree_alpha_stats_),
                _lib.ptr(self.tree_visit_cnt_), int(compact), int(subdivide), int(brute_force), _lib.ptr(nodes_o),
                _lib.ptr(w_o), _lib.ptr(a_o), cap, _lib.ptr(scratch), scratch_bytes, _lib.ptr(res),
                res.data_ptr() + 8, _lib.cur_stream()), "gf_octree_proc_device")
        host = res.cpu()
        n, err = int(host[:2].view(torch.int64).item()), int(host[2].item())
        if err & 1:
            raise RuntimeError("gf_octree_proc: the root was pruned (no valid leaf left in the octree)")
        if err & 2:
            raise RuntimeError("gf_octree_proc: a removed node is still linked (compact = 0 on a pruned tree?)")
        if err:
            raise RuntimeError(f"gf_octree_proc_device: error word {err}, {n} nodes needed, capacity {cap}")
        self.tree_nodes_gpu_ = nodes_o[:n * 128].clone() if n < cap else nodes_o
        self.tree_weight_stats_ = w_o[:n].clone() if n < cap else w_o
        self.tree_alpha_stats_ = a_o[:n].clone() if n < cap else a_o
        self.tree_visit_cnt_ = torch.zeros(n, dtype=torch.int64, device=dev)
        self._octree_stale = True          # the host mirror (self.octree) is refreshed when somebody asks for it

    def ProcOctreeHost(self, compact: bool, subdivide: bool, brute_force: bool):
        """The reference's own schedule of the same work: D2H of the node blob + stats, host compaction / subdivision
        in C++ (gf_octree_proc, csrc/octree_host.cu), H2D.  Kept as the checker of the device path."""
        import ctypes as C
        nodes = self.tree_nodes_gpu_.cpu().contiguous()
        w = self.tree_weight_stats_.cpu().contiguous()
        a = self.tree_alpha_stats_.cpu().contiguous()
        v = self.tree_visit_cnt_.cpu().contiguous()
        n_in, n_out, L = nodes.numel() // 128, C.c_int64(0), _lib.lib()
        args = (_lib.ptr(nodes), n_in, _lib.ptr(w), _lib.ptr(a), _lib.ptr(v), int(compact), int(subdivide),
                int(brute_force))
        _lib.check(L.gf_octree_proc(*args, None, None, None, 0, C.byref(n_out)), "gf_octree_proc")
        n = int(n_out.value)
        nodes_o = torch.empty(n * 128, dtype=torch.uint8)
        w_o, a_o = torch.empty(n, dtype=torch.int64), torch.empty(n, dtype=torch.int64)
        _lib.check(L.gf_octree_proc(*args, _lib.ptr(nodes_o), _lib.ptr(w_o), _lib.ptr(a_o), n, C.byref(n_out)),
                   "gf_octree_proc")
        oc = self._octree
        oc.load_blobs(nodes_o.numpy(), self.pers_trans_gpu_.cpu().numpy())
        oc.weight_stats, oc.alpha_stats = w_o.numpy(), a_o.numpy()
        oc.visit_cnt = np.zeros(n, np.int64)
        self._octree_stale = False
        self._upload_octree(stats=True)

    @property
    def octree(self) -> Optional[PersOctree]:
        """Host mirror of the device octree (numpy); downloaded again after the device changed the tree."""
        if getattr(self, "_octree_stale", False) and self._octree is not None:
            oc = self._octree
            oc.load_blobs(self.tree_nodes_gpu_.cpu().numpy(), self.pers_trans_gpu_.cpu().numpy())
            oc.weight_stats = self.tree_weight_stats_.cpu().numpy()
            oc.alpha_stats = self.tree_alpha_stats_.cpu().numpy()
            oc.visit_cnt = self.tree_visit_cnt_.cpu().numpy()
            self._octree_stale = False
        return self._octree

    @octree.setter
    def octree(self, value):
        self._octree = value
        self._octree_stale = False

    def MarkInvisibleNodes(self):
        """PersOctree::MarkInvisibleNodes (MarkInvisibleNodesKernel + CheckVisible, PersSampler_cuda.cu:680-742): a node
        no camera sees loses its transform.  One kernel over the device node blob (gf_octree_mark_invisible,
        csrc/octree_device.cu), the cameras staged through shared memory; same bits as the reference's kernel."""
        _lib.require_cuda(self.tree_nodes_gpu_, self.w2c_, self.intri_, self.bound_)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gf_octree_mark_invisible(
                _lib.ptr(self.tree_nodes_gpu_), self.n_nodes, _lib.ptr(self.w2c_), _lib.ptr(self.intri_),
                _lib.ptr(self.bound_), int(self.w2c_.shape[0]), _lib.cur_stream()), "gf_octree_mark_invisible")
        self._octree_stale = True

    def UpdateBlockIdxs(self, centers: torch.Tensor):
        """PersOctree::UpdateBlockIdxs (PersSampler_cuda.cu:746-798): nearest block centre per node
        (SetBlockIdxsNearestKernel -> gf_octree_set_block_idxs), then compact."""
        centers = torch.as_tensor(centers, dtype=torch.float32).to(self.device).contiguous().view(-1, 3)
        _lib.require_cuda(self.tree_nodes_gpu_)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gf_octree_set_block_idxs(
                _lib.ptr(self.tree_nodes_gpu_), self.n_nodes, _lib.ptr(centers), int(centers.shape[0]),
                _lib.cur_stream()), "gf_octree_set_block_idxs")
        self.ProcOctree(True, False, False)

    def UpdateMode(self, mode: int):
        self.mode_ = int(mode)

    def UpdateRayMarch(self, cur_step: int):
        """PersSampler.cpp:958-967 (fp32 arithmetic)."""
        if cur_step >= self.ray_march_fineness_decay_end_iter_:
            self.ray_march_fineness_ = 1.0
        else:
            progress = np.float32(cur_step) / np.float32(self.ray_march_fineness_decay_end_iter_)
            self.ray_march_fineness_ = float(np.exp(np.log(np.float32(1.)) * progress + np.log(
                np.float32(self.ray_march_init_fineness_)) * (np.float32(1.) - progress)))

    # ---- states (PersSampler.cpp:969-1016) ----------------------------------------------
    def States(self) -> List[torch.Tensor]:
        ms = torch.tensor(self.sub_div_milestones_, dtype=torch.int64, device=self.device)
        return [self.tree_nodes_gpu_, self.pers_trans_gpu_, self.tree_visit_cnt_, ms]

    def LoadStates(self, states: List[torch.Tensor], idx: int) -> int:
        self.tree_nodes_gpu_ = states[idx].clone().to(self.device).contiguous(); idx += 1
        self.pers_trans_gpu_ = states[idx].clone().to(self.device).contiguous(); idx += 1
        self.tree_visit_cnt_ = states[idx].clone().to(self.device).contiguous(); idx += 1
        self.sub_div_milestones_ = [int(v) for v in states[idx].cpu().tolist()]; idx += 1
        self._edge_pool = None
        self.octree.load_blobs(self.tree_nodes_gpu_.cpu().numpy(), self.pers_trans_gpu_.cpu().numpy())
        n = self.n_nodes
        self.tree_weight_stats_ = torch.full((n,), INIT_NODE_STAT, dtype=torch.int64, device=self.device)
        self.tree_alpha_stats_ = torch.full((n,), INIT_NODE_STAT, dtype=torch.int64, device=self.device)
        self.n_volumes_ = self.pers_trans_gpu_.numel() // 576
        return idx

    # ---- cold queries (bindings.cpp:42-299) ----------------------------------------------
    def TransQueryFrame(self, world_positions: torch.Tensor, anchors: torch.Tensor) -> torch.Tensor:
        """TransQueryFrameKernel (PersSampler_cuda.cu:854-922): anchors are TREE NODE indices."""
        wp = world_positions.contiguous().float()
        out = torch.zeros_like(wp)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gf_sampler_trans_query_frame(
                wp.shape[0], _lib.ptr(self.tree_nodes_gpu_), self.n_nodes, _lib.ptr(self.pers_trans_gpu_),
                _lib.ptr(anchors.contiguous().to(torch.int64)), _lib.ptr(wp), _lib.ptr(out), _lib.cur_stream()),
                "gf_sampler_trans_query_frame")
        return out

    def GetPointsAnchors(self, rays_origins, rays_dirs, t_starts, t_ends) -> torch.Tensor:
        """PersSampler::GetPointsAnchors (PersSampler_cuda.cu:924-980, proposal-sampler variant): i64 [R,S,1], the leaf
        containing each sample's mid-point t, -1 if none."""
        _lib.require_cuda(rays_origins, rays_dirs, t_starts, t_ends)
        R, S = t_starts.shape[0], t_starts.shape[1]
        t_cur = ((t_starts.float() + t_ends.float()) / 2.0).reshape(R, S).contiguous()     # :944
        anchors = torch.empty((R, S, 1), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gf_sampler_points_anchors(
                R, S, _lib.ptr(rays_origins.contiguous().float()), _lib.ptr(rays_dirs.contiguous().float()),
                _lib.ptr(t_cur), _lib.ptr(self.tree_nodes_gpu_), self.n_nodes, _lib.ptr(anchors), _lib.cur_stream()),
                "gf_sampler_points_anchors")
        return anchors

    def edge_pool(self) -> torch.Tensor:
        """PersOctree::ConstructEdgePool (PersSampler.cpp:833-893): uint8 [n_edges * 64] on the device.  The reference
        builds it once, in the PersOctree constructor, and keeps it while the octree is pruned and subdivided (the
        records name transforms, which are never removed); here it is built at the first use after InitSampler /
        LoadStates and kept the same way."""
        import ctypes as C
        cached = getattr(self, "_edge_pool", None)
        if cached is not None:
            return cached
        nodes = self.tree_nodes_gpu_.cpu().contiguous()
        n, L = C.c_int64(0), _lib.lib()
        _lib.check(L.gf_octree_edge_pool(_lib.ptr(nodes), self.n_nodes, None, 0, C.byref(n)), "gf_octree_edge_pool")
        pool = torch.empty(max(n.value, 1) * 64, dtype=torch.uint8)
        _lib.check(L.gf_octree_edge_pool(_lib.ptr(nodes), self.n_nodes, _lib.ptr(pool), n.value, C.byref(n)),
                   "gf_octree_edge_pool")
        self._edge_pool = pool[:n.value * 64].to(self.device)
        return self._edge_pool

    def GetEdgeSamples(self, n_pts: int, edge_idx: torch.Tensor = None, edge_coords: torch.Tensor = None):
        """PersSampler::GetEdgeSamples (PersSampler_cuda.cu:479-516): n_pts random points on faces shared by
        neighbouring leaves, warped by both leaves' transforms -> (f32 [n,2,3], i64 [n,2]).  edge_idx / edge_coords
        override the random draws (:498-499) for tests."""
        pool = self.edge_pool()
        n_edges = pool.numel() // 64
        if n_edges == 0:
            raise RuntimeError("GetEdgeSamples: the octree has no pair of neighbouring valid leaves")
        if edge_idx is None:
            edge_idx = torch.randint(0, n_edges, (n_pts,), device=self.device, dtype=torch.int64)
        if edge_coords is None:
            edge_coords = torch.rand((n_pts, 2), device=self.device) * 2.0 - 1.0
        out_pts = torch.empty((n_pts, 2, 3), dtype=torch.float32, device=self.device)
        out_idx = torch.empty((n_pts, 2), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gf_sampler_edge_samples(
                n_pts, _lib.ptr(pool), n_edges, _lib.ptr(self.pers_trans_gpu_), _lib.ptr(edge_idx.contiguous()),
                _lib.ptr(edge_coords.contiguous().float()), _lib.ptr(out_pts), _lib.ptr(out_idx), _lib.cur_stream()),
                "gf_sampler_edge_samples")
        return out_pts, out_idx

    def QueryTreeNodeCenters(self, anchors: torch.Tensor) -> torch.Tensor:
        center = self.tree_nodes_gpu_.view(-1, 128)[:, :12].contiguous().view(torch.float32)
        return center[anchors.to(torch.int64).clamp(0, self.n_nodes - 1)]

    def get_n_volumes(self) -> int:
        return self.n_volumes_

    def get_n_tree_nodes(self) -> int:
        return self.n_nodes

    def get_sampled_oct_per_ray(self) -> float:
        return self.sampled_oct_per_ray_

    def get_ray_march_fineness(self) -> float:
        return self.ray_march_fineness_

    def _node_field(self, lo, hi, dtype):
        return self.tree_nodes_gpu_.view(-1, 128)[:, lo:hi].contiguous().view(dtype)

    def get_tree_nodes_center_(self):
        return self._node_field(0, 12, torch.float32).cpu().tolist()

    def get_tree_nodes_side_len_(self):
        return self._node_field(12, 16, torch.float32).view(-1).cpu().tolist()

    def get_tree_nodes_trans_idx_(self):
        return self._node_field(96, 104, torch.int64).view(-1).cpu().tolist()

    def get_tree_nodes_block_idx_(self):
        return self._node_field(104, 112, torch.int64).view(-1).cpu().tolist()

    def get_tree_nodes_is_leaf_node_(self):
        return [bool(v) for v in self._node_field(88, 89, torch.uint8).view(-1).cpu().tolist()]


class PersSampler(nn.Module):
    """Drop-in for reference gfnerf/perssampler.py:47-447 (`generate_ray_samples` returns the same dense
    RaySamples structure; see rays.py)."""

    def __init__(self, c2w=None, intri=None, bounds=None, n_split_dataset=10, steps_per_split_dataset=10000,
                 steps_perssampler_init=30000, split_dist_thres: float = 1.5,
                 sub_div_milestones=(2000, 4000, 6000, 8000, 10000), compact_freq: int = 1000,
                 max_oct_intersect_per_ray: int = 1024, global_near: float = 0.01, scale_by_dis: bool = True,
                 bbox_levels: int = 8, sample_l: float = 1.0 / 256, max_level: int = 16, mode: int = 0,
                 sampled_oct_per_ray: int = 512, ray_march_fineness: float = 1.0, ray_march_init_fineness=16.0,
                 ray_march_fineness_decay_end_iter=10000, device=None, seed: int = 0, octree: PersOctree = None,
                 cameras=None):
        """Either (c2w [n,3,4], intri [n,3,3], bounds [n,2]) or, as the reference's call site does
        (gfnerf/nerfacto.py:223-227), `cameras=` an object with `camera_to_worlds` and `get_intrinsics_matrices()`
        (nerfstudio's Cameras / gfnerf_b200.Cameras) plus `bounds=` (perssampler.py:78-86)."""
        super().__init__()
        if cameras is not None:
            self.cameras = cameras
            c2w, intri = cameras.camera_to_worlds, cameras.get_intrinsics_matrices()
        if c2w is None or intri is None or bounds is None:
            raise TypeError("PersSampler: pass cameras= and bounds=, or c2w, intri and bounds")
        c2w = torch.as_tensor(c2w, dtype=torch.float32).detach().cpu()
        intri = torch.as_tensor(intri, dtype=torch.float32).detach().cpu()
        n = c2w.shape[0]
        w2c = torch.eye(4).unsqueeze(0).repeat(n, 1, 1)
        w2c[:, :3, :] = c2w
        w2c = torch.linalg.inv(w2c)[:, :3, :].contiguous()
        k = max(steps_perssampler_init // 30000, 1)
        self.max_pts_per_ray = MAX_SAMPLE_PER_RAY
        self.bounds = torch.as_tensor(bounds, dtype=torch.float32)
        self.sampler = PersSamplerCore()
        self.sampler.InitSampler(split_dist_thres, [int(x * k) for x in sub_div_milestones], compact_freq,
                                 max_oct_intersect_per_ray, global_near, scale_by_dis, bbox_levels, sample_l, max_level,
                                 c2w, w2c, torch.as_tensor(intri, dtype=torch.float32), self.bounds, mode,
                                 sampled_oct_per_ray, ray_march_fineness, ray_march_init_fineness,
                                 int(ray_march_fineness_decay_end_iter * k), device=device, seed=seed, octree=octree)
        self.n_split_dataset = n_split_dataset
        self.steps_per_split_dataset = steps_per_split_dataset
        self.steps_perssampler_init = steps_perssampler_init
        self.register_buffer("c2w", c2w.to(self.sampler.device))      # persistent, like perssampler.py:130
        self.cameras_labels = None    # int64 [n_cams,1] block label of every camera, set by the clustering (:236-237)
        self._register_state_dict_hook(self.state_dict_hook)

    def state_dict_hook(self, *args):
        destination, prefix = args[1], args[2]
        nodes, trans, visit, ms = self.sampler.States()
        destination[prefix + "tree_nodes_gpu"] = nodes
        destination[prefix + "pers_trans_gpu"] = trans
        destination[prefix + "tree_visit_cnt"] = visit
        destination[prefix + "milestones_ts"] = ms
        return destination

    def load_states(self, state_dict, prefix=""):
        """`state_dict` may also be the reference's positional form: a list of the four state tensors with `prefix`
        the start index (perssampler.py:579-580 `load_states(states, idx)`)."""
        if isinstance(state_dict, (list, tuple)):
            return self.sampler.LoadStates(list(state_dict), int(prefix or 0))
        pre = "" if prefix == "" else f"{prefix}."
        return self.sampler.LoadStates([state_dict[pre + k] for k in
                                        ("tree_nodes_gpu", "pers_trans_gpu", "tree_visit_cnt", "milestones_ts")], 0)

    def load_state_dict(self, state_dict, strict: bool = True):
        """perssampler.py:517-547: takes its four entries (and the cameras) out of the model's state dict."""
        self.load_states([state_dict.pop("persampler." + k) for k in
                          ("tree_nodes_gpu", "pers_trans_gpu", "tree_visit_cnt", "milestones_ts")], 0)
        # the reference reads `field.persampler.c2w` (:531); the same module is also reachable as `persampler`, so a
        # checkpoint carries the buffer under both names -- consume both, so that a strict load of the rest succeeds
        for key in ("persampler.c2w", "field.persampler.c2w"):
            if key in state_dict:
                self.c2w = state_dict.pop(key).to(self.sampler.device)
        return None

    # read-only views of the native sampler's configuration, under the reference's names (perssampler.py:583-623)
    sub_div_milestones_ = property(lambda self: list(self.sampler.sub_div_milestones_))
    compact_freq_ = property(lambda self: self.sampler.compact_freq_)
    max_oct_intersect_per_ray_ = property(lambda self: self.sampler.max_oct_intersect_per_ray_)
    global_near_ = property(lambda self: self.sampler.global_near_)
    sample_l_ = property(lambda self: self.sampler.sample_l_)
    scale_by_dis_ = property(lambda self: self.sampler.scale_by_dis_)
    mode_ = property(lambda self: self.sampler.mode_)
    n_volumes_ = property(lambda self: self.sampler.n_volumes_)
    sampled_oct_per_ray_ = property(lambda self: self.sampler.sampled_oct_per_ray_)
    ray_march_fineness_ = property(lambda self: self.sampler.ray_march_fineness_)

    # ---- eval-mode block routing (perssampler.py:138-165, 244-260): the block / appearance embedding of a render
    # chunk is the one of the training camera nearest to the chunk's first ray origin
    def get_nearest_split_dataset(self, origin, direction=None):
        dists = torch.linalg.norm(self.c2w[:, 0:3, -1] - origin.to(self.c2w.device), dim=1)
        nearest_image_idx = int(torch.argmin(dists).item())
        return int(self.cameras_labels.reshape(-1)[nearest_image_idx].item()), nearest_image_idx

    def get_nearest_split_dataset_orig(self, origin):
        n_images_per_split_dataset = self.c2w.shape[0] // self.n_split_dataset
        dists = torch.linalg.norm(self.c2w[:, 0:3, -1] - origin.to(self.c2w.device), dim=1)
        nearest_image_idx = int(torch.argmin(dists).item())
        cur_split_idx = min(nearest_image_idx // max(n_images_per_split_dataset, 1), self.n_split_dataset - 1)
        return cur_split_idx, nearest_image_idx

    def generate_ray_samples(self, ray_bundle):
        from .rays import Frustums, RaySamples, WarpedSamples
        rays_o, rays_d = ray_bundle.origins, ray_bundle.directions
        S = self.max_pts_per_ray
        cur_step, cur_split_idx, nearest_image_idx = -1, -1, None
        if ray_bundle.steps is not None:
            cur_step = int(ray_bundle.steps.reshape(-1)[0].item())
            if cur_step >= self.steps_perssampler_init:
                cur_split_idx = ((cur_step - self.steps_perssampler_init) // self.steps_per_split_dataset) % self.n_split_dataset
        if cur_split_idx == -1 and rays_o.shape[0] > 0:       # eval mode or init stage (:367-376)
            if self.cameras_labels is not None:
                cur_split_idx, nearest_image_idx = self.get_nearest_split_dataset(rays_o[0])
            else:
                cur_split_idx, nearest_image_idx = self.get_nearest_split_dataset_orig(rays_o[0])
        bounds = self.bounds[0, :].to(rays_o.device).repeat((rays_o.shape[0], 1))
        world, warp, dirs, dists, ts, anchors, start_end, first = self.sampler.GetSamples(rays_o, rays_d, bounds)
        f2 = WarpedSamples(sampled_world_pts=world, sampled_pts=warp, sampled_dirs=dirs, sampled_dists=dists.unsqueeze(-1),
                           sampled_t=ts.unsqueeze(-1), sampled_anchors=anchors,
                           pts_idx_start_end=start_end.unsqueeze(1).expand(-1, S, -1),
                           first_oct_dis=first.unsqueeze(1).expand(-1, S, -1))
        fr = Frustums(origins=rays_o.unsqueeze(1).expand(-1, S, -1), directions=rays_d.unsqueeze(1).expand(-1, S, -1),
                      starts=ts.unsqueeze(-1), ends=ts.unsqueeze(-1),
                      pixel_area=None if ray_bundle.pixel_area is None else ray_bundle.pixel_area.unsqueeze(1).expand(-1, S, -1))
        cam = None if ray_bundle.camera_indices is None else ray_bundle.camera_indices.unsqueeze(1).expand(-1, S, -1)
        if ray_bundle.rel_camera_indices is not None:
            rel = ray_bundle.rel_camera_indices.unsqueeze(1).expand(-1, S, -1)
        elif cam is not None and nearest_image_idx is not None:   # test mode: the nearest training image's embedding (:429-432)
            rel = torch.ones_like(cam) * nearest_image_idx
        else:
            rel = None
        return RaySamples(f2samples=f2, frustums=fr, camera_indices=cam, rel_camera_indices=rel,
                          deltas=dists.unsqueeze(-1), cur_step=cur_step, cur_split_dataset_idx=cur_split_idx)

    forward = generate_ray_samples

    def update_oct_nodes(self, sampled_anchors, pts_idx_bounds, sampled_weights, sampled_alpha, iter_step):
        self.sampler.UpdateOctNodes(sampled_anchors, pts_idx_bounds, sampled_weights, sampled_alpha, iter_step)

    def update_ray_march(self, cur_step: int):
        self.sampler.UpdateRayMarch(cur_step)

    def update_mode(self, mode: int):
        self.sampler.UpdateMode(mode)

    def update_block_idx(self, block_centers):
        self.sampler.UpdateBlockIdxs(block_centers)

    def trans_query_frame(self, world_positions_flat, anchors_flat):
        return self.sampler.TransQueryFrame(world_positions_flat, anchors_flat)

    def query_tree_nodes_centers(self, anchors):
        return self.sampler.QueryTreeNodeCenters(anchors)

    def get_points_anchors(self, rays_o, rays_d, t_starts, t_ends):
        return self.sampler.GetPointsAnchors(rays_o, rays_d, t_starts, t_ends)

    def get_edge_samples(self, n_pts: int):
        return self.sampler.GetEdgeSamples(n_pts)

    def states(self):
        return self.sampler.States()

    def get_n_volumes(self):
        return self.sampler.get_n_volumes()
