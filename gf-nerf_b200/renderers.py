"""RGB / depth / accumulation renderers with the call signatures of the reference
(nerfstudio/model_components/renderers.py: `RGBRenderer` :58-138, `AccumulationRenderer` :195-221,
`DepthRenderer` :224-283) for callers that hold per-sample weights.

They consume the weights `RaySamples.get_weights_f2nerf` returns; the per-ray sums are plain reductions over the
sample axis.  The fused training / render path (`engine.GFNeRFEngine`) never materialises weights for them: it gets
rgb, depth and accumulation out of the same warp-per-ray kernel that scans the transmittance.
"""
import torch
from torch import nn


class RGBRenderer(nn.Module):
    def __init__(self, background_color="random") -> None:
        super().__init__()
        self.background_color = background_color

    @classmethod
    def combine_rgb(cls, rgb, weights, background_color="random", ray_indices=None, num_rays=None):
        """sum_s w_s c_s; the reference's background blend is commented out (renderers.py:109)."""
        if ray_indices is not None:
            raise NotImplementedError("packed samples (nerfacc) are not used by gf-nerf")
        return torch.sum(weights * rgb, dim=-2)

    def forward(self, rgb, weights, ray_indices=None, num_rays=None):
        if not self.training:
            rgb = torch.nan_to_num(rgb)
        out = self.combine_rgb(rgb, weights, background_color=self.background_color, ray_indices=ray_indices,
                               num_rays=num_rays)
        if not self.training:
            torch.clamp_(out, min=0.0, max=1.0)
        return out


class AccumulationRenderer(nn.Module):
    @classmethod
    def forward(cls, weights, ray_indices=None, num_rays=None):
        if ray_indices is not None:
            raise NotImplementedError("packed samples (nerfacc) are not used by gf-nerf")
        return torch.sum(weights, dim=-2)


class DepthRenderer(nn.Module):
    def __init__(self, method="median") -> None:
        super().__init__()
        self.method = method

    def forward(self, weights, ray_samples, ray_indices=None, num_rays=None):
        steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
        if self.method == "median":
            cumulative = torch.cumsum(weights[..., 0], dim=-1)
            split = torch.ones((*weights.shape[:-2], 1), device=weights.device) * 0.5
            idx = torch.clamp(torch.searchsorted(cumulative, split, side="left"), 0, steps.shape[-2] - 1)
            return torch.gather(steps[..., 0], dim=-1, index=idx)
        if self.method == "expected":
            depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + 1e-10)
            # global clip; the dense tensor's min is the padding's 0 (renderers.py:281)
            return torch.clip(depth, steps.min(), steps.max())
        raise NotImplementedError(f"Method {self.method} not implemented")
