#!/usr/bin/env python
"""Generates tests/golden/ref_*.npz by RUNNING THE REFERENCE's own Python code for the pieces of the
per-ray hot path that import without its native extension:

  nerfstudio/cameras/rays.py:178-200           RaySamples.get_weights_f2nerf
  nerfstudio/model_components/renderers.py     RGBRenderer.combine_rgb / forward, DepthRenderer('expected'),
                                               AccumulationRenderer
  nerfstudio/field_components/activations.py   trunc_exp (+ its custom backward)
  nerfstudio/model_components/losses.py:73-84  CharbonnierLoss
  gfnerf/mlp.py:3-57                           MLPNetwork (the two stacks of gfnerf/nerfacto_field.py:174-179,217-227)
  torch.optim.Adam                             as configured at gfnerf/config.py:132-135 (lr 1e-2, eps 1e-15)

Run in the build container (needs /root/reference; it does not exist on the GPU box, so the vectors are
committed):

  python tests/golden/make_golden.py

`torchtyping` and `nerfacc` are not installed here; two empty stand-ins (type annotations only / an import
that the dense path never calls) are put on sys.path.  The tcnn SH encoding is NOT importable (un-vendored
third party): the head's SH input is the oracle's restatement of it (oracle.sh4) evaluated on stored unit
directions, so the fixture pins the MLP given that encoding and parity for SH itself stays unpinned
(SURVEY.md 8c).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get("GF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    tt = types.ModuleType("torchtyping")

    class TensorType:
        def __class_getitem__(cls, item):
            return cls

    tt.TensorType = TensorType
    tt.patch_typeguard = lambda: None
    sys.modules["torchtyping"] = tt
    sys.modules["nerfacc"] = types.ModuleType("nerfacc")
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("gf_ref_mlp", os.path.join(REF, "gfnerf", "mlp.py"))
    mlp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mlp)
    from nerfstudio.cameras.rays import Frustums, RaySamples
    from nerfstudio.field_components.activations import trunc_exp
    from nerfstudio.model_components import renderers
    from nerfstudio.model_components.losses import CharbonnierLoss
    return mlp.MLPNetwork, Frustums, RaySamples, trunc_exp, renderers, CharbonnierLoss


def ragged_counts(rng, R, S):
    counts = rng.randint(0, S + 1, size=R)
    counts[0] = 0          # a ray that hit nothing
    counts[1] = S          # a full ray
    counts[2] = 1
    return counts.astype(np.int32)


def composite_case(Frustums, RaySamples, renderers, R=48, S=96, seed=3):
    """Dense, zero-padded [R,S,.] tensors exactly as PersSampler.generate_ray_samples hands them over
    (padding has delta = 0, t = 0), through get_weights_f2nerf and the three renderers, with autograd."""
    rng = np.random.RandomState(seed)
    counts = ragged_counts(rng, R, S)
    mask = (np.arange(S)[None, :] < counts[:, None])
    sigma = (rng.gamma(0.6, 8.0, size=(R, S)) * mask).astype(np.float32)
    sigma[3, : counts[3]] = 0.0                    # an empty-space ray
    sigma[4, : counts[4]] *= 1e3                   # an opaque ray (transmittance underflows)
    delta = (rng.uniform(0.002, 0.02, size=(R, S)) * mask).astype(np.float32)
    t = (np.cumsum(rng.uniform(0.01, 0.05, size=(R, S)), axis=1) * mask).astype(np.float32)
    rgb = (rng.uniform(0, 1, size=(R, S, 3)) * mask[..., None]).astype(np.float32)
    g_rgb = rng.normal(size=(R, 3)).astype(np.float32)
    g_acc = rng.normal(size=(R, 1)).astype(np.float32)

    dens = torch.tensor(sigma[..., None], requires_grad=True)
    col = torch.tensor(rgb, requires_grad=True)
    tt = torch.tensor(t[..., None])
    z = torch.zeros(R, S, 3)
    fr = Frustums(origins=z, directions=z, starts=tt, ends=tt, pixel_area=torch.zeros(R, S, 1))
    rs = RaySamples(frustums=fr, deltas=torch.tensor(delta[..., None]))
    w, a, T = rs.get_weights_f2nerf(dens)
    rgb_r = renderers.RGBRenderer(background_color="last_sample")
    rgb_r.train()
    out_rgb = rgb_r(rgb=col, weights=w)
    acc = renderers.AccumulationRenderer()(weights=w)
    depth = renderers.DepthRenderer(method="expected")(weights=w, ray_samples=rs)
    (out_rgb * torch.tensor(g_rgb)).sum().add((acc * torch.tensor(g_acc)).sum()).backward()
    rgb_r.eval()
    with torch.no_grad():
        out_rgb_eval = rgb_r(rgb=col, weights=w)
    return dict(counts=counts, sigma=sigma, delta=delta, t=t, rgb=rgb, g_rgb=g_rgb, g_acc=g_acc[:, 0],
                weights=w.detach().numpy()[..., 0], alphas=a.detach().numpy()[..., 0],
                trans=T.detach().numpy()[..., 0], out_rgb=out_rgb.detach().numpy(),
                out_rgb_eval=out_rgb_eval.numpy(), out_acc=acc.detach().numpy()[:, 0],
                out_depth=depth.detach().numpy()[:, 0], d_sigma=dens.grad.numpy()[..., 0], d_rgb=col.grad.numpy())


def mlp_case(MLPNetwork, trunc_exp, n=384, R=12, H=64, seed=5):
    """GFNeRFField.get_density + get_outputs (gfnerf/nerfacto_field.py:455,491-503,540-555) on explicit inputs:
    h = base(feat); density = trunc_exp(h[:,0:1] + 1); rgb = head(cat[SH(16), h[:,1:16], emb(32)])."""
    torch.manual_seed(seed)
    rng = np.random.RandomState(seed)
    cfg = lambda out_act: {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": out_act,
                           "n_neurons": H, "n_hidden_layers": 1}
    base = MLPNetwork(32, 16, cfg("None"))
    head_cfg = cfg("Sigmoid")
    head_cfg["n_hidden_layers"] = 2
    head = MLPNetwork(63, 3, head_cfg)
    lin = [m for m in list(base.modules()) + list(head.modules()) if isinstance(m, torch.nn.Linear)]
    assert [tuple(l.weight.shape) for l in lin] == [(H, 32), (16, H), (H, 63), (H, H), (3, H)], \
        [tuple(l.weight.shape) for l in lin]
    params = np.concatenate([np.concatenate([l.weight.detach().numpy().ravel(), l.bias.detach().numpy().ravel()])
                             for l in lin]).astype(np.float32)
    # hash features are fp16 values (Hash3DAnchored_cuda.cu:195 widens fp16 output)
    feat = rng.uniform(-0.05, 0.05, size=(n, 32)).astype(np.float16).astype(np.float32)
    ray_id = np.sort(rng.randint(0, R, size=n)).astype(np.int32)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import oracle as orc
    dirs = rng.normal(size=(R, 3))
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    sh = orc.sh4(dirs)                             # stands in for tcnn SH (fp16-rounded values)
    emb = rng.normal(size=(R, 32)).astype(np.float32)
    g_sigma = rng.normal(size=(n, 1)).astype(np.float32) * 0.1
    g_rgb = rng.normal(size=(n, 3)).astype(np.float32) * 0.1

    x = torch.tensor(feat, requires_grad=True)
    e = torch.tensor(emb, requires_grad=True)
    h = base(x)
    density = trunc_exp(h[:, 0:1] + 1.0)
    idx = torch.tensor(ray_id, dtype=torch.long)
    rgb = head(torch.cat([torch.tensor(sh)[idx], h[:, 1:16], e[idx]], dim=-1))
    ((density * torch.tensor(g_sigma)).sum() + (rgb * torch.tensor(g_rgb)).sum()).backward()
    d_params = np.concatenate([np.concatenate([l.weight.grad.numpy().ravel(), l.bias.grad.numpy().ravel()])
                               for l in lin]).astype(np.float32)
    return dict(H=np.int32(H), params=params, feat=feat, ray_id=ray_id, dirs=dirs, sh=sh, emb=emb, g_sigma=g_sigma[:, 0],
                g_rgb=g_rgb, sigma=density.detach().numpy()[:, 0], rgb=rgb.detach().numpy(),
                d_feat=x.grad.numpy(), d_params=d_params, d_emb=e.grad.numpy())


def loss_adam_case(CharbonnierLoss, R=200, n=4096, seed=7):
    rng = np.random.RandomState(seed)
    pred = rng.uniform(0, 1, size=(R, 3)).astype(np.float32)
    target = rng.uniform(0, 1, size=(R, 3)).astype(np.float32)
    target[:5] = pred[:5]                         # zero residuals: sqrt(eps^2) branch
    p = torch.tensor(pred, requires_grad=True)
    loss = CharbonnierLoss()(p, torch.tensor(target))
    loss.backward()
    # Adam, the optimizer of every gf-nerf parameter group (gfnerf/config.py:132-135): three steps
    w0 = rng.uniform(-1e-2, 1e-2, size=n).astype(np.float32)
    grads = (rng.normal(size=(3, n)) * np.array([1e-3, 1.0, 1e-6])[:, None]).astype(np.float32)
    grads[:, :64] = 0.0                           # untouched table rows keep moving with their momentum
    w = torch.nn.Parameter(torch.tensor(w0))
    opt = torch.optim.Adam([w], lr=1e-2, eps=1e-15)
    traj = []
    for g in grads:
        w.grad = torch.tensor(g)
        opt.step()
        traj.append(w.detach().numpy().copy())
    return dict(pred=pred, target=target, loss=np.float32(loss.item()), g_pred=p.grad.numpy(), adam_w0=w0,
                adam_grads=grads, adam_traj=np.stack(traj))


def s3im_case(R=512, seed=9):
    """S3IM (nerfstudio/model_components/losses.py:713-794) as gf-nerf configures it (kernel 4, stride 4, repeat 10,
    patch height 32, gfnerf/nerfacto.py:186-197).  Its random re-patching draws torch.randperm from the global CPU
    generator; the same seed replayed gives the index list stored in the fixture."""
    from nerfstudio.model_components.losses import S3IM
    rng = np.random.RandomState(seed)
    src = rng.uniform(0, 1, size=(R, 3)).astype(np.float32)
    tar = np.clip(src + rng.normal(0, 0.1, size=(R, 3)), 0, 1).astype(np.float32)
    loss_fn = S3IM(s3im_kernel_size=4, s3im_stride=4, s3im_repeat_time=10, s3im_patch_height=32)
    x = torch.tensor(src, requires_grad=True)
    torch.manual_seed(1234)
    loss = loss_fn(x, torch.tensor(tar))
    loss.backward()
    torch.manual_seed(1234)
    index = torch.cat([torch.arange(R)] + [torch.randperm(R) for _ in range(9)]).numpy().astype(np.int64)
    return dict(src=src, tar=tar, index=index, loss=np.float64(loss.item()), g_src=x.grad.numpy())


def scheduler_case():
    """GFNerfExponentialDecayScheduler.get_scheduler (nerfstudio/engine/schedulers.py:138-184) through torch's
    LambdaLR, for the gf-nerf init stage and a block-stage configuration."""
    # nerfstudio/configs/base_config.py does not import under Python 3.12 (a mutable dataclass default at :118); the
    # scheduler only needs its InstantiateConfig base, so that one class is stood in and schedulers.py itself is the
    # reference's file, unmodified
    import dataclasses
    import importlib.util
    stub = types.ModuleType("nerfstudio.configs.base_config")

    @dataclasses.dataclass
    class InstantiateConfig:
        _target: type

        def setup(self, **kwargs):
            return self._target(self, **kwargs)

    stub.InstantiateConfig = InstantiateConfig
    sys.modules["nerfstudio.configs.base_config"] = stub
    spec = importlib.util.spec_from_file_location("gf_ref_schedulers",
                                                  os.path.join(REF, "nerfstudio", "engine", "schedulers.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["gf_ref_schedulers"] = mod
    spec.loader.exec_module(mod)
    GFNerfExponentialDecayScheduler = mod.GFNerfExponentialDecayScheduler
    GFNerfExponentialDecaySchedulerConfig = mod.GFNerfExponentialDecaySchedulerConfig
    out = {}
    cfgs = {
        "init": dict(lr_final=1e-4, max_steps=30000, warmup_steps=0, steps_perssampler_init=30000,
                     steps_per_split_dataset=10000, n_split_dataset=10),
        "block": dict(lr_final=5e-4, max_steps=10000, warmup_steps=100, ramp="linear", lr_pre_warmup=1e-6,
                      steps_perssampler_init=300, steps_per_split_dataset=200, n_split_dataset=3),
    }
    steps = np.array([0, 1, 50, 99, 100, 299, 300, 301, 499, 500, 899, 900, 901, 1500, 5000, 29999, 30000, 45000])
    for name, kw in cfgs.items():
        cfg = GFNerfExponentialDecaySchedulerConfig(**kw)
        w = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([w], lr=1e-2)
        sched = GFNerfExponentialDecayScheduler(cfg).get_scheduler(opt, 1e-2)
        mult = np.array([sched.lr_lambdas[0](int(s)) for s in steps], np.float64)
        out[name + "_mult"] = mult
        out[name + "_cfg"] = np.array(repr(kw))
    out["steps"] = steps
    return out


def main():
    MLPNetwork, Frustums, RaySamples, trunc_exp, renderers, CharbonnierLoss = _import_reference()
    torch.set_num_threads(1)
    out = {
        "ref_composite": composite_case(Frustums, RaySamples, renderers),
        "ref_mlp": mlp_case(MLPNetwork, trunc_exp),
        "ref_loss_adam": loss_adam_case(CharbonnierLoss),
        "ref_scheduler": scheduler_case(),
        "ref_s3im": s3im_case(),
    }
    for name, d in out.items():
        path = os.path.join(HERE, name + ".npz")
        with tempfile.NamedTemporaryFile(dir=HERE, suffix=".npz", delete=False) as f:
            np.savez_compressed(f, **d)
        os.replace(f.name, path)
        print(path, {k: getattr(v, "shape", ()) for k, v in d.items()})


if __name__ == "__main__":
    main()
