"""Ray / sample containers and `get_weights_f2nerf` -- the nerfstudio-side operator API of the path.

Mirrors reference nerfstudio/cameras/rays.py: `Frustums` (:29-106), `WarpedSamples` (:108-117),
`RaySamples` (:126-200, incl. `get_weights_f2nerf` :178-200) and `RayBundle` (:224-262), with the same
field names so that `GFNeRFModel.get_outputs` (gfnerf/nerfacto.py:522-619) reads them unchanged.  They are plain
dataclasses (the reference's `TensorDataclass` broadcasting machinery is caller-side and out of scope).

`get_weights_f2nerf` runs the warp-per-ray scan kernel (csrc/composite.cu) on the dense `[R,S,1]` layout --
every ray is a row of S slots whose padding has delta = 0 -- with a `torch.autograd.Function` whose backward is
`gf_composite_backward` fed with the gradient flowing into the weights.  CUDA tensors only.
"""
from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional

import torch

from . import _lib


@dataclass
class Frustums:
    origins: torch.Tensor
    directions: torch.Tensor
    starts: torch.Tensor
    ends: torch.Tensor
    pixel_area: Optional[torch.Tensor] = None
    offsets: Optional[torch.Tensor] = None

    @property
    def shape(self):
        return self.starts.shape[:-1]

    def get_positions(self) -> torch.Tensor:
        """nerfstudio/cameras/rays.py:46-56"""
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        if self.offsets is not None:
            pos = pos + self.offsets
        return pos


@dataclass
class WarpedSamples:
    sampled_world_pts: torch.Tensor
    sampled_pts: torch.Tensor
    sampled_dirs: torch.Tensor
    sampled_dists: torch.Tensor
    sampled_t: torch.Tensor
    sampled_anchors: torch.Tensor
    pts_idx_start_end: torch.Tensor
    first_oct_dis: torch.Tensor


class _WeightsF2NeRF(torch.autograd.Function):
    """(weights, alphas, transmittance) = f(densities) for fixed deltas; only `weights` carries gradient
    (alphas / transmittance feed the octree vote, gfnerf/nerfacto.py:607-613)."""

    @staticmethod
    def forward(ctx, densities, deltas):
        _lib.require_cuda(densities, deltas)
        shape = densities.shape
        R, S = int(shape[0]), int(shape[-2])
        sigma = densities.detach().reshape(-1).contiguous().float()
        delta = deltas.detach().reshape(-1).contiguous().float()
        offsets = torch.arange(R + 1, device=sigma.device, dtype=torch.int32) * S
        w, a, t = (torch.empty_like(sigma) for _ in range(3))
        with torch.cuda.device(sigma.device):
            _lib.check(_lib.lib().gf_composite_forward(
                R, _lib.ptr(offsets), _lib.ptr(sigma), _lib.ptr(delta), None, None, _lib.ptr(w), _lib.ptr(a),
                _lib.ptr(t), None, None, None, None, _lib.cur_stream()), "gf_composite_forward")
        ctx.save_for_backward(offsets, sigma, delta, t)
        ctx.shape = shape
        ctx.mark_non_differentiable(a, t)
        return w.view(shape), a.view(shape), t.view(shape)

    @staticmethod
    def backward(ctx, g_w, _ga, _gt):
        offsets, sigma, delta, trans = ctx.saved_tensors
        g_w = g_w.reshape(-1).contiguous().float()
        d_sigma = torch.empty_like(sigma)
        with torch.cuda.device(sigma.device):
            _lib.check(_lib.lib().gf_composite_backward(
                offsets.numel() - 1, _lib.ptr(offsets), _lib.ptr(sigma), _lib.ptr(delta), None, _lib.ptr(trans),
                None, None, _lib.ptr(g_w), _lib.ptr(d_sigma), None, _lib.cur_stream()), "gf_composite_backward")
        return d_sigma.view(ctx.shape), None


@dataclass
class RaySamples:
    frustums: Frustums
    f2samples: Optional[WarpedSamples] = None
    camera_indices: Optional[torch.Tensor] = None
    rel_camera_indices: Optional[torch.Tensor] = None
    deltas: Optional[torch.Tensor] = None
    spacing_starts: Optional[torch.Tensor] = None
    spacing_ends: Optional[torch.Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Optional[Dict[str, Any]] = None
    times: Optional[torch.Tensor] = None
    cur_step: Any = None
    cur_split_dataset_idx: Any = None

    @property
    def shape(self):
        return self.frustums.shape

    def get_weights_f2nerf(self, densities: torch.Tensor):
        """alphas = 1 - exp(-delta sigma); T = exp(-exclusive cumsum(delta sigma)); weights = alphas T
        (nerfstudio/cameras/rays.py:178-200).  Returns (weights, alphas, transmittance), each [R,S,1]."""
        return _WeightsF2NeRF.apply(densities, self.deltas)

    def get_weights(self, densities: torch.Tensor) -> torch.Tensor:
        return self.get_weights_f2nerf(densities)[0]


@dataclass
class RayBundle:
    origins: torch.Tensor
    directions: torch.Tensor
    lookat_directions: Optional[torch.Tensor] = None
    pixel_area: Optional[torch.Tensor] = None
    camera_indices: Optional[torch.Tensor] = None
    rel_camera_indices: Optional[torch.Tensor] = None
    nears: Optional[torch.Tensor] = None
    fars: Optional[torch.Tensor] = None
    metadata: Optional[Dict[str, Any]] = None
    times: Optional[torch.Tensor] = None
    steps: Optional[torch.Tensor] = None

    def __len__(self):
        return int(self.origins.shape[0])
