"""`GFNeRFModel.get_outputs` through the operator API (reference gfnerf/nerfacto.py:522-619): sampler ->
field -> get_weights_f2nerf -> RGB / depth / accumulation renderers -> octree feedback, on the reference's dense
`[R,1024,...]` tensors.  This is the drop-in path for callers that keep nerfstudio's Trainer / autograd / optimizers;
`engine.GFNeRFEngine` is the same computation fused end to end on the compact layout (what bench.py times).
"""
from typing import Dict

import torch
from torch import nn

from .field import FieldHeadNames, GFNeRFField
from .perssampler import PersSampler
from .rays import RayBundle
from .renderers import AccumulationRenderer, DepthRenderer, RGBRenderer


class GFNeRFModel(nn.Module):
    def __init__(self, persampler: PersSampler, field: GFNeRFField, scale_factor: float = 1.0,
                 background_color="last_sample"):
        super().__init__()
        self.persampler, self.field, self.scale_factor = persampler, field, float(scale_factor)
        self.field.persampler = persampler
        self.renderer_rgb = RGBRenderer(background_color=background_color)
        self.renderer_accumulation = AccumulationRenderer()
        self.renderer_depth = DepthRenderer(method="expected")     # gfnerf/nerfacto.py:288

    def get_outputs(self, ray_bundle: RayBundle) -> Dict[str, torch.Tensor]:
        ray_samples = self.persampler(ray_bundle)
        field_outputs = self.field(ray_samples)
        weights, alphas, trans = ray_samples.get_weights_f2nerf(field_outputs[FieldHeadNames.DENSITY])
        rgb = self.renderer_rgb(rgb=field_outputs[FieldHeadNames.RGB], weights=weights)
        depth = self.renderer_depth(weights=weights, ray_samples=ray_samples) / self.scale_factor
        accumulation = self.renderer_accumulation(weights=weights)
        oct_depth = ray_samples.f2samples.first_oct_dis[:, 0, :] / self.scale_factor
        outputs = {"rgb": rgb, "accumulation": accumulation, "depth": depth, "oct_depth": oct_depth}
        if self.training and ray_bundle.steps is not None:          # nerfacto.py:598-616
            cur_step = int(ray_bundle.steps.reshape(-1)[0].item())
            self.persampler.update_ray_march(cur_step)
            self.persampler.update_mode(0)
            if self.field.cur_stage == "init_stage":
                self.persampler.update_oct_nodes(sampled_anchors=ray_samples.f2samples.sampled_anchors,
                                                 pts_idx_bounds=ray_samples.f2samples.pts_idx_start_end,
                                                 sampled_weights=weights.detach(), sampled_alpha=alphas,
                                                 iter_step=cur_step)
            else:
                self.persampler.update_mode(1)
        return outputs

    forward = get_outputs
