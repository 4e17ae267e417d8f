#!/usr/bin/env python
"""Generates tests/golden/ref_field.npz by RUNNING THE REFERENCE's own field class -- gfnerf/nerfacto_field.py
GFNeRFField.get_density / get_outputs (:412-591), the caller of the Hash3DAnchored encoding and of the two MLPs --
unmodified, on the CPU, with stand-ins only for what the reference itself loads from outside its Python tree:

  * `torch.classes.my_classes.Hash3DAnchored` (its native extension, gfnerf/hash_3d_anchored.py:13-25): a fake whose
    AnchoredQuery runs the reference's OWN forward kernel compiled for the host (oracle/_ref, oracle/ref_host.py);
  * `tinycudann.Encoding` (un-vendored third party): SphericalHarmonics degree 4 = the oracle's restatement of tcnn's
    SH (fp16 output, like tcnn); the Frequency encoding the field constructs but never calls = a dummy;
  * `torchtyping`, `nerfacc` (not installed): empty modules; `dataclasses.dataclass` tolerates the dataclass-instance
    defaults of nerfstudio/configs/base_config.py that Python 3.12 rejects.

So the fixture pins the WIRING of SURVEY rows a13 / a15 to the reference's code: (pts + 1.5) / 3, anchor column 0,
the validity mask, base_network, the 1 + 15 split, trunc_exp(h0 + 1), (dir + 1) / 2 -> SH, the embedding lookup by
rel_camera_indices, the concatenation order [SH 16 | geo 15 | appearance 32], mlp_head -- with autograd gradients of
every MLP / embedding parameter.  (Finding while writing it: get_density hard-codes `self.cur_stage = 'init_stage'`
at :449, so in this snapshot of the reference the block-stage residual branch below it is unreachable.)

  make -C oracle ref && python tests/golden/make_golden_field.py          # build container only
"""
import dataclasses
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from oracle import ref_host as rh  # noqa: E402
from tests.helpers import fast_primes  # noqa: E402


class FakeHash3DAnchored:
    """Stands in for torch.classes.my_classes.Hash3DAnchored (hashanchored/bindings.cpp:300-357)."""
    instances = []

    def __init__(self, log2_table_size, n_volumes, lr):
        self.local = 1 << log2_table_size
        self.n_volumes = n_volumes
        rng = np.random.RandomState(11)
        self.feat = torch.tensor(rng.uniform(-0.3, 0.3, size=(16 * self.local, 2)).astype(np.float32))
        self.prim = fast_primes(16 * n_volumes * 3, 5).reshape(16, n_volumes, 3)
        self.bias = np.zeros((16 * n_volumes, 3), np.float32)
        FakeHash3DAnchored.instances.append(self)

    def AnchoredQuery(self, points, anchors):
        out = rh.hash_forward(self.feat.numpy(), self.prim, self.bias, points.detach().numpy(), anchors.numpy(), "fma")
        self.last = (points.detach().numpy().copy(), anchors.numpy().copy(), out.copy())
        return torch.from_numpy(out)

    def Reset(self):
        pass                                       # keep the table drawn above (Reset would re-draw U(-0.01, 0.01))

    def Zero(self):
        self.feat.zero_()

    def States(self):
        return [self.feat, torch.from_numpy(self.prim), torch.from_numpy(self.bias), torch.full((1,), self.n_volumes, dtype=torch.int32)]

    def GetParams(self):
        return [self.feat]

    def SetFeatPoolRequireGrad(self, flag):
        pass

    def to(self, device):
        pass

    def ReleaseResources(self):
        pass


class FakeTcnnEncoding(torch.nn.Module):
    """tinycudann.Encoding for the two configs the field constructs (nerfacto_field.py:152-166)."""

    def __init__(self, n_input_dims, encoding_config):
        super().__init__()
        self.otype = encoding_config["otype"]
        self.n_output_dims = 16 if self.otype == "SphericalHarmonics" else n_input_dims * 2 * encoding_config["n_frequencies"]

    def forward(self, x):
        assert self.otype == "SphericalHarmonics"
        d = (x.detach().numpy().astype(np.float32) * np.float32(2.0) - np.float32(1.0)).astype(np.float32)
        return torch.from_numpy(orc.sh4(d)).half()          # tcnn returns its preferred precision, fp16


def import_reference_field():
    tt = types.ModuleType("torchtyping")

    class TensorType:
        def __class_getitem__(cls, item):
            return cls

    tt.TensorType = TensorType
    tt.patch_typeguard = lambda: None
    sys.modules.setdefault("torchtyping", tt)
    sys.modules.setdefault("nerfacc", types.ModuleType("nerfacc"))
    tcnn = types.ModuleType("tinycudann")
    tcnn.Encoding = FakeTcnnEncoding
    sys.modules["tinycudann"] = tcnn
    torch.classes.load_library = lambda path: None           # "../GF-NeRF/gfnerf/bindings/f2nerf-bindings.so"
    torch.classes.my_classes = types.SimpleNamespace(Hash3DAnchored=FakeHash3DAnchored)
    orig = dataclasses.dataclass

    def tolerant(cls=None, /, **kw):
        def wrap(c):
            try:
                return orig(c, **kw)
            except ValueError as e:
                if "mutable default" not in str(e):
                    raise
                for val in list(vars(c).values()):
                    if dataclasses.is_dataclass(val) and type(val).__hash__ is None:
                        type(val).__hash__ = object.__hash__
                return orig(c, **kw)
        return wrap if cls is None else wrap(cls)

    dataclasses.dataclass = tolerant
    if REF not in sys.path:
        sys.path.insert(0, REF)
    try:
        import importlib
        field_mod = importlib.import_module("gfnerf.nerfacto_field")
        rays = importlib.import_module("nerfstudio.cameras.rays")
    finally:
        dataclasses.dataclass = orig
    return field_mod, rays


def main():
    field_mod, rays = import_reference_field()
    torch.manual_seed(3)
    R, S, n_vol, n_img, log2T = 24, 40, 6, 5, 10
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        field = field_mod.GFNeRFField(aabb=torch.tensor([[-1., -1., -1.], [1., 1., 1.]]), num_images=n_img,
                                      log2_hashmap_size=log2T, use_appearance_embedding=True, n_volumes=n_vol,
                                      base_dir=tmp, steps_perssampler_init=10000)
    field.train()
    rng = np.random.RandomState(0)
    counts = rng.randint(0, S + 1, size=R)
    counts[0], counts[1] = 0, S
    m = np.arange(S)[None, :] < counts[:, None]
    warp = (rng.uniform(-1.2, 1.2, size=(R, S, 3)) * m[..., None]).astype(np.float32)         # zero padding, like the sampler's
    anchors = np.zeros((R, S, 3), np.int64)
    anchors[..., 0] = rng.randint(0, n_vol, size=(R, S)) * m
    anchors[..., 1] = rng.randint(0, 50, size=(R, S)) * m
    dirs = rng.normal(size=(R, 3))
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    cam = rng.randint(0, n_img, size=R).astype(np.int64)
    T = torch.from_numpy
    dirs_rs = T(dirs)[:, None, :].expand(R, S, 3).contiguous()
    z1, z3 = torch.zeros(R, S, 1), torch.zeros(R, S, 3)
    f2 = rays.WarpedSamples(sampled_world_pts=z3, sampled_pts=T(warp), sampled_dirs=dirs_rs, sampled_dists=z1,
                            sampled_t=z1, sampled_anchors=T(anchors), pts_idx_start_end=torch.zeros(R, S, 2),
                            first_oct_dis=z1)
    fr = rays.Frustums(origins=z3, directions=dirs_rs, starts=z1, ends=z1, pixel_area=z1)
    rs = rays.RaySamples(frustums=fr, f2samples=f2, camera_indices=T(cam)[:, None, None].expand(R, S, 1).contiguous(),
                         rel_camera_indices=T(cam)[:, None, None].expand(R, S, 1).contiguous(),
                         cur_step=torch.zeros(R, S, 1), cur_split_dataset_idx=torch.full((R, S, 1), -1.0))
    density, geo = field.get_density(rs)
    out = field.get_outputs(rs, density_embedding=geo)
    rgb = out[field_mod.FieldHeadNames.RGB]
    g_sigma = (rng.normal(size=(R, S, 1)) * 0.1).astype(np.float32)
    g_rgb = (rng.normal(size=(R, S, 3)) * 0.1).astype(np.float32)
    ((density * T(g_sigma)).sum() + (rgb * T(g_rgb)).sum()).backward()
    lin = [mod for mod in list(field.base_network.modules()) + list(field.mlp_head.modules())
           if isinstance(mod, torch.nn.Linear)]
    assert [tuple(l.weight.shape) for l in lin] == [(64, 32), (16, 64), (64, 63), (64, 64), (3, 64)]
    flat = lambda get: np.concatenate([np.concatenate([get(l.weight).numpy().ravel(), get(l.bias).numpy().ravel()])
                                       for l in lin]).astype(np.float32)
    hash3d = FakeHash3DAnchored.instances[0]
    fx = dict(warp_pts=warp, anchors=anchors, counts=counts.astype(np.int32), dirs=dirs, cam=cam, g_sigma=g_sigma,
              g_rgb=g_rgb, table=hash3d.feat.numpy(), prim=hash3d.prim, log2T=np.int64(log2T),
              params=flat(lambda p: p.detach()), d_params=flat(lambda p: p.grad),
              emb=field.embedding_appearance.embedding.weight.detach().numpy(),
              d_emb=field.embedding_appearance.embedding.weight.grad.numpy(),
              density=density.detach().numpy(), rgb=rgb.detach().numpy(), geo=geo.detach().numpy(),
              query_pts=hash3d.last[0], query_anchors=hash3d.last[1], hash_feats=hash3d.last[2].astype(np.float16))
    path = os.path.join(HERE, "ref_field.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB; density range", float(density.min()), float(density.max()),
          "queried points", hash3d.last[0].shape[0], "of", R * S)


if __name__ == "__main__":
    main()
