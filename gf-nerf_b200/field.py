"""`GFNeRFField`: the nerfstudio `Field` of GF-NeRF over the fused kernels -- operator API of the path.

Mirrors reference gfnerf/nerfacto_field.py: constructor arguments and sub-module names (:98-246), `get_density`
+ `get_outputs` as one `forward` (:412-591; `Field.forward`, nerfstudio/fields/base_field.py:108-129, is the only
caller, gfnerf/nerfacto.py:535), the per-block residual tables of the focal stage (`add_table / del_table /
save_table / load_table`, :248-403; residual added to the hash features before the frozen MLP, :458-489), and
`parameters()` (:593-603).  State-dict keys match the reference (`base_network.layers.N.*`, `mlp_head.layers.N.*`,
`embedding_appearance.embedding.weight`, `base_encoding_init.{feat_pool,prime_pool,bias_pool,n_volumes}`).

What differs, by design: the field evaluates only the VALID samples of the dense `[R,1024]` layout (the reference
pushes all 1024 padded slots of every ray through hash + both MLPs; padding has delta = 0 and therefore weight 0),
so padded slots of the returned density / rgb are 0 instead of the reference's don't-care values; the two MLP
stacks run as one fused tensor-core kernel; SH and the appearance embedding enter per ray.
Gradients reach `feat_pool`, the nn.Linear parameters and the embedding through one `torch.autograd.Function`,
so any torch optimizer (the reference: Adam, gfnerf/config.py:132-135) trains it.
"""
import os
from enum import Enum
from pathlib import Path
from typing import Dict, Iterator, Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from .hash_3d_anchored import Hash3DAnchored
from .mlp import MLPNetwork
from .rays import RaySamples

HASH_DIM = 32


class FieldHeadNames(Enum):
    """nerfstudio/field_components/field_heads.py:27-38 (the two heads gf-nerf produces)"""
    RGB = "rgb"
    DENSITY = "density"


class Embedding(nn.Module):
    """nerfstudio/field_components/embedding.py:25-50"""

    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.embedding = nn.Embedding(in_dim, out_dim)

    def forward(self, in_tensor):
        return self.embedding(in_tensor)


def _scale(n_rays: int) -> float:
    """power-of-two loss scale for the fp16 gradient fragments of the MLP backward (dL/drgb ~ 1/R)"""
    return float(2 ** int(np.ceil(np.log2(max(n_rays, 1)))))


class _FusedFieldFn(torch.autograd.Function):
    """(sigma [V], rgb [V,3]) = field(pts01 [V,3], anchors [V], ray_id [V]; per-ray dirs, embeddings)."""

    @staticmethod
    def forward(ctx, feat_pool, res_pool, mlp_blob, ray_emb, enc, res_enc, pts01, anchors, ray_id, dirs, hidden,
                train_mlp):
        L, st = _lib.lib(), _lib.cur_stream()
        _lib.require_cuda(pts01, anchors, ray_id, dirs, mlp_blob)
        V, R = pts01.shape[0], dirs.shape[0]
        dev = pts01.device
        feat = torch.empty((V, HASH_DIM), dtype=torch.float16, device=dev)
        enc.launch_forward(pts01, anchors, out_f16=feat)
        if res_enc is not None:   # focal stage: residual at the hash-feature level (nerfacto_field.py:477-489)
            res_enc.launch_forward_residual(pts01, anchors, feat)
        blob = mlp_blob.detach().contiguous().float()
        emb = None if ray_emb is None else ray_emb.detach().contiguous().float()
        ray_bias = torch.empty((R, hidden), dtype=torch.float32, device=dev)
        sigma = torch.empty(V, dtype=torch.float32, device=dev)
        rgb = torch.empty((V, 3), dtype=torch.float32, device=dev)
        # the forward's ReLU masks, handed to the backward (include/gfnerf_b200.h gf_mlp_forward)
        masks = torch.empty((V, int(L.gf_mlp_mask_words(hidden))), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.gf_mlp_ray_bias(R, hidden, _lib.ptr(blob), _lib.ptr(dirs), _lib.ptr(emb), _lib.ptr(ray_bias),
                                         st), "gf_mlp_ray_bias")
            _lib.check(L.gf_mlp_forward(V, None, hidden, _lib.ptr(blob), _lib.ptr(feat), _lib.ptr(ray_id),
                                        _lib.ptr(ray_bias), _lib.ptr(sigma), _lib.ptr(rgb), _lib.ptr(masks), st),
                       "gf_mlp_forward")
        ctx.save_for_backward(pts01, anchors, ray_id, dirs, feat, blob, ray_bias, emb if emb is not None else dirs,
                              masks)
        ctx.meta = (enc, res_enc, hidden, bool(train_mlp), emb is not None, feat_pool.shape,
                    None if res_pool is None else res_pool.shape)
        return sigma, rgb

    @staticmethod
    def backward(ctx, d_sigma, d_rgb):
        L, st = _lib.lib(), _lib.cur_stream()
        pts01, anchors, ray_id, dirs, feat, blob, ray_bias, emb, masks = ctx.saved_tensors
        enc, res_enc, hidden, train_mlp, has_emb, pool_shape, res_shape = ctx.meta
        emb = emb if has_emb else None
        V, R, dev = pts01.shape[0], dirs.shape[0], pts01.device
        d_sigma = d_sigma.contiguous().float()
        d_rgb = d_rgb.contiguous().float()
        d_feat = torch.empty((V, HASH_DIM), dtype=torch.float16, device=dev)
        d_blob = torch.zeros_like(blob) if train_mlp else None
        d_rb = torch.zeros((R, hidden), dtype=torch.float32, device=dev) if train_mlp else None
        d_emb = torch.zeros((R, 32), dtype=torch.float32, device=dev) if (train_mlp and has_emb) else None
        with torch.cuda.device(dev):
            _lib.check(L.gf_mlp_backward(V, None, hidden, _lib.ptr(blob), _lib.ptr(feat), _lib.ptr(ray_id),
                                         _lib.ptr(ray_bias), _lib.ptr(masks), _lib.ptr(d_sigma), _lib.ptr(d_rgb),
                                         _lib.ptr(d_feat),
                                         _lib.ptr(d_blob), _lib.ptr(d_rb), _scale(R), st), "gf_mlp_backward")
            if train_mlp:
                _lib.check(L.gf_mlp_ray_bias_backward(R, hidden, _lib.ptr(blob), _lib.ptr(dirs), _lib.ptr(emb),
                                                      _lib.ptr(d_rb), _lib.ptr(d_blob), _lib.ptr(d_emb), st),
                           "gf_mlp_ray_bias_backward")
        g_pool = g_res = None
        if ctx.needs_input_grad[0]:
            g_pool = torch.zeros(pool_shape, dtype=torch.float32, device=dev)
            enc.launch_backward(pts01, anchors, d_feat, True, g_pool)
        if res_enc is not None and ctx.needs_input_grad[1]:
            g_res = torch.zeros(res_shape, dtype=torch.float32, device=dev)
            res_enc.launch_backward(pts01, anchors, d_feat, True, g_res)
        return g_pool, g_res, d_blob, d_emb, None, None, None, None, None, None, None, None


class GFNeRFField(nn.Module):
    def __init__(self, aabb, num_images: int, num_layers: int = 2, hidden_dim: int = 64, geo_feat_dim: int = 15,
                 num_levels: int = 16, max_res: int = 2048, log2_hashmap_size: int = 19, num_layers_color: int = 3,
                 num_layers_transient: int = 2, hidden_dim_color: int = 64, hidden_dim_transient: int = 64,
                 appearance_embedding_dim: int = 32, transient_embedding_dim: int = 16,
                 use_transient_embedding: bool = False, use_semantics: bool = False, num_semantic_classes: int = 100,
                 pass_semantic_gradients: bool = False, use_pred_normals: bool = False,
                 use_average_appearance_embedding: bool = False, use_appearance_embedding: bool = False,
                 spatial_distortion=None, n_blocks: int = 1, n_active_block: int = 3,
                 steps_perssampler_init: int = 10000, block_centers=None, base_dir: str = "", n_volumes: int = 0,
                 generator: torch.Generator = None) -> None:
        super().__init__()
        if not (num_layers == 2 and num_layers_color == 3 and hidden_dim == hidden_dim_color and geo_feat_dim == 15
                and appearance_embedding_dim == 32 and num_levels == 16):
            raise ValueError("gfnerf_b200 GFNeRFField: the fused kernel is built for the gf-nerf field shape "
                             "(32 -> H -> 16, 63 -> H -> H -> 3)")
        if _lib.lib().gf_mlp_param_count(hidden_dim) < 0:
            raise ValueError(f"gfnerf_b200 GFNeRFField: hidden width {hidden_dim} is not built")
        if use_semantics or use_pred_normals or use_transient_embedding:
            raise NotImplementedError("semantics / predicted normals / transient heads are not used by gf-nerf")
        self.register_buffer("aabb", torch.as_tensor(aabb))
        self.geo_feat_dim, self.hidden_dim = geo_feat_dim, hidden_dim
        self.num_images = num_images
        self.appearance_embedding_dim = appearance_embedding_dim
        self.use_appearance_embedding = use_appearance_embedding
        self.embedding_appearance = Embedding(num_images, appearance_embedding_dim)
        cfg = lambda out_act, n_hidden: {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": out_act,
                                         "n_neurons": hidden_dim, "n_hidden_layers": n_hidden}
        self.base_network = MLPNetwork(32, 1 + geo_feat_dim, cfg("None", num_layers - 1))
        self.mlp_head = MLPNetwork(16 + geo_feat_dim + appearance_embedding_dim, 3, cfg("Sigmoid", num_layers_color - 1))
        self.base_dir = base_dir
        self.encodings_ckpt_dir = Path(base_dir) / "encodings_ckpt"
        self.n_volumes, self.log2_table_size = n_volumes, log2_hashmap_size
        self.n_blocks, self.block_centers, self.n_active_block = n_blocks, block_centers, n_active_block
        self.steps_perssampler_init = steps_perssampler_init
        self.single_mlp = True
        self.base_encoding_init = Hash3DAnchored(log2_hashmap_size, n_volumes, generator=generator)
        self.base_encoding_init.reset()                       # U(-0.01, 0.01), nerfacto_field.py:200
        self.active_block_idxs, self.active_block_idxs_test = [], []
        self.cur_stage, self.cur_step, self.cur_split_dataset_idx = "init_stage", -1, -1
        self.persampler = None

    # ---- focal-stage tables (nerfacto_field.py:248-403) ------------------------------------------
    def add_table(self, table_idx: int, generator: torch.Generator = None) -> None:
        """zero-initialised residual encoder `base_encoding_{table_idx}` (nerfacto_field.py:336-347)"""
        enc = Hash3DAnchored(self.log2_table_size, self.n_volumes, generator=generator)
        enc.zero()
        setattr(self, f"base_encoding_{table_idx}", enc)

    def del_table(self, table_idx: int) -> None:
        enc = getattr(self, f"base_encoding_{table_idx}")
        enc.unregister_hooks()
        enc.release_resources()
        delattr(self, f"base_encoding_{table_idx}")

    def save_table(self, table_idx: int) -> str:
        """nerfacto_field.py:368-383: the encoder's state dict (feat_pool / prime_pool / bias_pool / n_volumes, no
        prefix) under encodings_ckpt/base_encoding_{idx}.ckpt -- the reference's file format, so swapped-out tables
        travel between the two implementations."""
        os.makedirs(self.encodings_ckpt_dir, exist_ok=True)
        path = str(self.encodings_ckpt_dir / f"base_encoding_{table_idx}.ckpt")
        enc = getattr(self, f"base_encoding_{table_idx}")
        torch.save({k: v.detach().cpu() for k, v in enc.state_dict().items()}, path)
        return path

    def load_table(self, table_idx: int, strict: bool = False) -> None:
        """nerfacto_field.py:386-403; a missing file is an error only if `strict`.  Unlike the reference the table is
        created if it is not there yet (the reference asserts it is)."""
        path = str(self.encodings_ckpt_dir / f"base_encoding_{table_idx}.ckpt")
        if not os.path.exists(path):
            if strict:
                raise FileNotFoundError(path)
            return
        if not hasattr(self, f"base_encoding_{table_idx}"):
            self.add_table(table_idx)
        states = torch.load(path, map_location="cpu")
        enc = getattr(self, f"base_encoding_{table_idx}")
        if isinstance(states, (list, tuple)):      # files written before the format followed the reference's
            enc.load_states(list(states), 0)
        else:
            enc.load_state_dict(dict(states))

    def update_active_blocks(self, cur_split_data_idx: int, ray_samples=None) -> None:
        """nerfacto_field.py:248-330: make block `cur_split_data_idx` the active one for the current mode (train /
        eval): bring its table in (from disk if it was swapped out), freeze the eval-side tables, and swap every table
        that is neither active in training nor in eval out to disk.  -1 = no active block."""
        active = [] if cur_split_data_idx == -1 else [int(cur_split_data_idx)]
        in_train, in_test = (active, self.active_block_idxs_test) if self.training else (self.active_block_idxs, active)
        for i in in_test:
            if not hasattr(self, f"base_encoding_{i}"):
                self.add_table(i)
                self.load_table(i, strict=True)
            getattr(self, f"base_encoding_{i}").set_require_grad(False)
        for i in in_train:
            if not hasattr(self, f"base_encoding_{i}"):
                self.add_table(i)
                self.load_table(i, strict=False)
            enc = getattr(self, f"base_encoding_{i}")
            enc.set_require_grad(True)
            enc.train()
        for i in sorted(set(range(self.n_blocks)) - set(in_train) - set(in_test)):
            if hasattr(self, f"base_encoding_{i}"):
                self.save_table(i)
                self.del_table(i)
        if self.training:
            self.active_block_idxs = active
        else:
            self.active_block_idxs_test = active

    def memory_stats(self):
        print(f"Allocated Memory: {torch.cuda.memory_allocated() / 1024 ** 3:.2f} GB")
        print(f"Cached Memory: {torch.cuda.memory_reserved() / 1024 ** 3:.2f} GB")

    def set_stage(self, stage: str, active_block: Optional[int] = None) -> None:
        """'init_stage' (global table + MLPs + embedding train) or 'block_stage' (they are frozen, the active
        block's residual table trains; nerfacto_field.py:458-489, 530-549)."""
        assert stage in ("init_stage", "block_stage")
        self.cur_stage = stage
        init = stage == "init_stage"
        self.base_encoding_init.set_require_grad(init)
        self.base_network.requires_grad_(init)
        self.mlp_head.requires_grad_(init)
        self.embedding_appearance.embedding.requires_grad_(init)
        self.active_block_idxs = [] if init else [int(active_block)]
        self.active_block_idxs_test = list(self.active_block_idxs)

    # ---- forward ---------------------------------------------------------------------------
    def forward(self, ray_samples: RaySamples, compute_normals: bool = False) -> Dict[FieldHeadNames, torch.Tensor]:
        if compute_normals:
            raise NotImplementedError("normals are not used by gf-nerf (config.predict_normals is False)")
        f2 = ray_samples.f2samples
        R, S = f2.sampled_pts.shape[0], f2.sampled_pts.shape[1]
        dev = f2.sampled_pts.device
        se = f2.pts_idx_start_end[:, 0, :]
        counts = (se[:, 1] - se[:, 0]).to(torch.int64)
        valid = (torch.arange(S, device=dev)[None, :] < counts[:, None]).reshape(-1)
        idx = valid.nonzero(as_tuple=False).squeeze(1)                           # one host sync (the reference: five)
        pts01 = ((f2.sampled_pts.reshape(-1, 3)[idx] + 1.5) / 3.0).contiguous()   # nerfacto_field.py:431
        anchors = f2.sampled_anchors.reshape(-1, f2.sampled_anchors.shape[-1])[idx, 0].contiguous()
        ray_id = (idx // S).to(torch.int32).contiguous()
        dirs = ray_samples.frustums.directions[:, 0, :].contiguous().float()
        ray_emb = None
        if self.use_appearance_embedding:
            if ray_samples.rel_camera_indices is None:
                raise AttributeError("Camera indices are not provided.")
            rel = ray_samples.rel_camera_indices.reshape(R, -1)[:, 0].to(torch.int64)
            ray_emb = self.embedding_appearance(rel)
        blob = torch.cat([self.base_network.flat_params(), self.mlp_head.flat_params()])
        res_enc = res_pool = None
        if self.cur_stage == "block_stage":
            active = self.active_block_idxs if self.training else self.active_block_idxs_test
            assert len(active) == 1                                            # nerfacto_field.py:476
            res = getattr(self, f"base_encoding_{active[0]}")
            res_enc, res_pool = res.hash_3d, res.hash_3d.feat_pool_
        sigma, rgb = _FusedFieldFn.apply(self.base_encoding_init.hash_3d.feat_pool_, res_pool, blob, ray_emb,
                                         self.base_encoding_init.hash_3d, res_enc, pts01, anchors, ray_id, dirs,
                                         self.hidden_dim, self.cur_stage == "init_stage")
        density = torch.zeros(R * S, dtype=torch.float32, device=dev).index_put((idx,), sigma).view(R, S, 1)
        color = torch.zeros((R * S, 3), dtype=torch.float32, device=dev).index_put((idx,), rgb).view(R, S, 3)
        return {FieldHeadNames.DENSITY: density, FieldHeadNames.RGB: color}

    def get_density(self, ray_samples: RaySamples):
        """Density alone (the geo features stay inside the fused kernel, so the second return value of the
        reference, base_mlp_out, is not materialised)."""
        return self.forward(ray_samples)[FieldHeadNames.DENSITY], None

    def get_outputs(self, ray_samples: RaySamples, density_embedding=None):
        return {FieldHeadNames.RGB: self.forward(ray_samples)[FieldHeadNames.RGB]}

    def parameters(self, recurse: bool = True) -> Iterator[nn.Parameter]:
        """nerfacto_field.py:593-603: registered parameters + the global table (block tables get their own
        optimizer, gfnerf/nerfacto.py:478-488)."""
        seen = set()
        for p in super().parameters(recurse):
            seen.add(id(p))
            yield p
        for p in self.base_encoding_init.parameters(recurse):
            if id(p) not in seen:
                yield p
