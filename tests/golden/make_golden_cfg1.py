#!/usr/bin/env python
"""Generates tests/golden/ref_cfg1.npz by RUNNING THE REFERENCE's own CPU-capable torch path -- BASELINE.json
configs[0] ("nerfstudio nerfacto with torch hash-encoding backend, 4096 synthetic rays x 48 samples, fwd+bwd on
CPU"), the path the north star asks to be timed next to the GPU numbers -- at a size small enough to commit.

The reference classes are imported unmodified from /root/reference and composed exactly as TorchNerfactoField does
(nerfstudio/fields/nerfacto_field.py:370-461; `nerfstudio.fields.*` itself does not import under Python 3.12,
BASELINE.md section 3):

  nerfstudio/field_components/encodings.py:220-354   HashEncoding(implementation="torch")
  nerfstudio/field_components/encodings.py           SHEncoding(levels=4)  (utils/math.py:27-87)
  nerfstudio/field_components/mlp.py:25-97           MLP (base 3 x 64 ReLU out, head 2 x 32 ReLU out)
  nerfstudio/field_components/field_heads.py:96-117  DensityFieldHead (softplus), RGBFieldHead (sigmoid)
  nerfstudio/field_components/embedding.py           Embedding(num_images, 40)
  nerfstudio/cameras/rays.py:155-177                 RaySamples.get_weights
  nerfstudio/model_components/renderers.py           RGBRenderer / AccumulationRenderer / DepthRenderer("expected")

The fixture holds the inputs, every parameter the reference initialised (so that the restatement
oracle/nerfacto_cpu.py runs on identical numbers), the rendered outputs, the MSE loss and the gradients.

  python tests/golden/make_golden_cfg1.py            # needs /root/reference (build container only)
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    tt = types.ModuleType("torchtyping")

    class TensorType:
        def __class_getitem__(cls, item):
            return cls

    tt.TensorType = TensorType
    tt.patch_typeguard = lambda: None
    sys.modules.setdefault("torchtyping", tt)
    sys.modules.setdefault("nerfacc", types.ModuleType("nerfacc"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from nerfstudio.cameras.rays import Frustums, RaySamples
    from nerfstudio.field_components.embedding import Embedding
    from nerfstudio.field_components.encodings import HashEncoding, SHEncoding
    from nerfstudio.field_components.field_heads import DensityFieldHead, RGBFieldHead
    from nerfstudio.field_components.mlp import MLP
    from nerfstudio.model_components import renderers
    return dict(Frustums=Frustums, RaySamples=RaySamples, Embedding=Embedding, HashEncoding=HashEncoding,
                SHEncoding=SHEncoding, DensityFieldHead=DensityFieldHead, RGBFieldHead=RGBFieldHead, MLP=MLP,
                renderers=renderers)


def cfg1_inputs(R, S, n_images, seed):
    """positions U[0,1)^3, unit directions (one per ray), deltas U(0,0.05), target U[0,1)."""
    rng = np.random.RandomState(seed)
    pos = rng.uniform(0, 1, size=(R, S, 3)).astype(np.float32)
    d = rng.normal(size=(R, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    delta = rng.uniform(0, 0.05, size=(R, S, 1)).astype(np.float32)
    ends = np.cumsum(delta, axis=1).astype(np.float32)
    starts = (ends - delta).astype(np.float32)
    cam = rng.randint(0, n_images, size=(R,)).astype(np.int64)
    target = rng.uniform(0, 1, size=(R, 3)).astype(np.float32)
    return dict(pos=pos, dirs=d, delta=delta, starts=starts, ends=ends, cam=cam, target=target)


class ReferenceCfg1(torch.nn.Module):
    """The reference's components wired like TorchNerfactoField.get_density / get_outputs."""

    def __init__(self, ref, log2_hashmap_size=19, n_images=16):
        super().__init__()
        self.ref = ref
        self.position_encoding = ref["HashEncoding"](num_levels=16, min_res=16, max_res=2048,
                                                     log2_hashmap_size=log2_hashmap_size, features_per_level=2,
                                                     implementation="torch")
        self.direction_encoding = ref["SHEncoding"](levels=4)
        self.mlp_base = ref["MLP"](in_dim=32, num_layers=3, layer_width=64, skip_connections=(4,),
                                   out_activation=torch.nn.ReLU())
        self.embedding_appearance = ref["Embedding"](n_images, 40)
        self.mlp_head = ref["MLP"](in_dim=64 + 16 + 40, num_layers=2, layer_width=32, out_activation=torch.nn.ReLU())
        self.field_output_density = ref["DensityFieldHead"](in_dim=64)
        self.field_head_rgb = ref["RGBFieldHead"](in_dim=32)
        self.rgb_renderer = ref["renderers"].RGBRenderer(background_color="last_sample")
        self.acc_renderer = ref["renderers"].AccumulationRenderer()
        self.depth_renderer = ref["renderers"].DepthRenderer(method="expected")

    def forward(self, inp):
        R, S = inp["pos"].shape[:2]
        pos = torch.as_tensor(inp["pos"])
        dirs = torch.as_tensor(inp["dirs"])[:, None, :].expand(R, S, 3)
        fr = self.ref["Frustums"](origins=torch.zeros(R, S, 3), directions=dirs, starts=torch.as_tensor(inp["starts"]),
                                  ends=torch.as_tensor(inp["ends"]), pixel_area=torch.zeros(R, S, 1))
        rs = self.ref["RaySamples"](frustums=fr, deltas=torch.as_tensor(inp["delta"]))
        base_out = self.mlp_base(self.position_encoding(pos))
        density = self.field_output_density(base_out)
        emb = self.embedding_appearance(torch.as_tensor(inp["cam"]))[:, None, :].expand(R, S, 40)
        h = self.mlp_head(torch.cat([self.direction_encoding(dirs), base_out, emb], dim=-1))
        rgb = self.field_head_rgb(h)
        w = rs.get_weights(density)
        out = dict(rgb=self.rgb_renderer(rgb=rgb, weights=w), accumulation=self.acc_renderer(weights=w),
                   depth=self.depth_renderer(weights=w, ray_samples=rs), weights=w, density=density, sample_rgb=rgb)
        out["loss"] = torch.nn.functional.mse_loss(out["rgb"], torch.as_tensor(inp["target"]))
        return out

    def linears(self):
        return ([*self.mlp_base.layers] + [self.field_output_density.net] + [*self.mlp_head.layers]
                + [self.field_head_rgb.net])

    def export_params(self):
        p = {"hash_table": self.position_encoding.hash_table.detach().numpy().copy(),
             "scalings": self.position_encoding.scalings.numpy().copy(),
             "embedding": self.embedding_appearance.embedding.weight.detach().numpy().copy()}
        for i, l in enumerate(self.linears()):
            p[f"w{i}"] = l.weight.detach().numpy().copy()
            p[f"b{i}"] = l.bias.detach().numpy().copy()
        return p

    def export_grads(self):
        g = {"g_hash_table": self.position_encoding.hash_table.grad.numpy().copy(),
             "g_embedding": self.embedding_appearance.embedding.weight.grad.numpy().copy()}
        for i, l in enumerate(self.linears()):
            g[f"g_w{i}"] = l.weight.grad.numpy().copy()
            g[f"g_b{i}"] = l.bias.grad.numpy().copy()
        return g


def main():
    ref = import_reference()
    torch.manual_seed(11)
    R, S, n_images, log2 = 96, 48, 16, 11
    model = ReferenceCfg1(ref, log2_hashmap_size=log2, n_images=n_images)
    model.train()
    # the reference's 1e-3 table init leaves every density at softplus(~0): scale the table up so that weights,
    # transmittance and the table gradient are exercised over a useful range
    with torch.no_grad():
        model.position_encoding.hash_table.mul_(300.0)
    inp = cfg1_inputs(R, S, n_images, seed=1234)
    out = model(inp)
    out["loss"].backward()
    fx = {f"in_{k}": v for k, v in inp.items()}
    fx.update({f"p_{k}": v for k, v in model.export_params().items()})
    fx.update(model.export_grads())
    for k in ("rgb", "accumulation", "depth", "weights", "density", "sample_rgb", "loss"):
        fx[f"out_{k}"] = out[k].detach().numpy()
    fx["log2_hashmap_size"] = np.int64(log2)
    path = os.path.join(HERE, "ref_cfg1.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB; loss", float(out["loss"]),
          "acc range", float(out["accumulation"].min()), float(out["accumulation"].max()))


if __name__ == "__main__":
    main()
