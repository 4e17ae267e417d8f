"""GPU parity against golden vectors made by RUNNING the reference's own Python code
(tests/golden/make_golden.py): composite + renderers (1e-5, fp32), the field MLP (1e-2, the fp16 MLP class),
Charbonnier and Adam (1e-5 / 2e-6).  Everything goes through the C-ABI (gfnerf_b200._lib)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: d[k] for k in d.files}


def T(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(a, b, rtol, what):
    a, b = a.detach().cpu().numpy().astype(np.float64), np.asarray(b, np.float64)
    scale = max(float(np.abs(b).max()), 1e-30)
    err = float(np.abs(a - b).max())
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


def test_composite_matches_reference_renderers():
    from gfnerf_b200 import _lib
    L, st = _lib.lib(), _lib.cur_stream()
    g = load("ref_composite")
    counts = g["counts"]
    R, S = g["sigma"].shape
    m = np.arange(S)[None, :] < counts[:, None]
    offsets = T(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))
    sigma, delta, rgb, t = T(g["sigma"][m]), T(g["delta"][m]), T(g["rgb"][m]), T(g["t"][m])
    V = sigma.numel()
    w, a, tr = (torch.empty(V, device="cuda") for _ in range(3))
    out_rgb, depth, acc = torch.empty((R, 3), device="cuda"), torch.empty(R, device="cuda"), torch.empty(R, device="cuda")
    tmax = torch.zeros(1, device="cuda")
    _lib.check(L.gf_composite_forward(R, _lib.ptr(offsets), _lib.ptr(sigma), _lib.ptr(delta), _lib.ptr(rgb), _lib.ptr(t),
                                      _lib.ptr(w), _lib.ptr(a), _lib.ptr(tr), _lib.ptr(out_rgb), _lib.ptr(depth),
                                      _lib.ptr(acc), _lib.ptr(tmax), st))
    close(w, g["weights"][m], 1e-5, "weights")
    close(a, g["alphas"][m], 1e-5, "alphas")
    close(tr, g["trans"][m], 1e-5, "transmittance")
    close(out_rgb, g["out_rgb"], 1e-5, "rgb")
    close(acc, g["out_acc"], 1e-5, "accumulation")
    close(torch.minimum(depth.clamp_min(0.0), tmax), g["out_depth"], 1e-5, "depth")
    d_sigma, d_rgb = torch.empty(V, device="cuda"), torch.empty((V, 3), device="cuda")
    g_rgb, g_acc = T(g["g_rgb"]), T(g["g_acc"])     # keep the tensors alive across the call
    _lib.check(L.gf_composite_backward(R, _lib.ptr(offsets), _lib.ptr(sigma), _lib.ptr(delta), _lib.ptr(rgb),
                                       _lib.ptr(tr), _lib.ptr(g_rgb), _lib.ptr(g_acc), None,
                                       _lib.ptr(d_sigma), _lib.ptr(d_rgb), st))
    close(d_sigma, g["d_sigma"][m], 2e-5, "d_sigma")
    close(d_rgb, g["d_rgb"][m], 1e-5, "d_rgb")


def test_mlp_matches_reference_mlpnetwork():
    from gfnerf_b200 import _lib
    L, st = _lib.lib(), _lib.cur_stream()
    g = load("ref_mlp")
    H, n, R = int(g["H"]), g["feat"].shape[0], g["dirs"].shape[0]
    p, feat, ray_id, dirs, emb = T(g["params"]), T(g["feat"].astype(np.float16)), T(g["ray_id"]), T(g["dirs"]), T(g["emb"])
    rb = torch.empty((R, H), device="cuda")
    _lib.check(L.gf_mlp_ray_bias(R, H, _lib.ptr(p), _lib.ptr(dirs), _lib.ptr(emb), _lib.ptr(rb), st))
    sigma, rgb = torch.empty(n, device="cuda"), torch.empty((n, 3), device="cuda")
    masks = torch.zeros((n, 2, 4), dtype=torch.int32, device="cuda")
    _lib.check(L.gf_mlp_forward(n, None, H, _lib.ptr(p), _lib.ptr(feat), _lib.ptr(ray_id), _lib.ptr(rb),
                                _lib.ptr(sigma), _lib.ptr(rgb), _lib.ptr(masks), st))
    close(sigma, g["sigma"], 2e-3, "density")     # north star: 1e-2 for the fp16 MLP; split precision does better
    close(rgb, g["rgb"], 2e-3, "rgb")
    d_feat = torch.empty((n, 32), dtype=torch.float16, device="cuda")
    d_params = torch.zeros(g["params"].size, device="cuda")
    d_rb = torch.zeros((R, H), device="cuda")
    d_emb = torch.zeros((R, 32), device="cuda")
    g_sigma, g_rgb = T(g["g_sigma"]), T(g["g_rgb"])
    _lib.check(L.gf_mlp_backward(n, None, H, _lib.ptr(p), _lib.ptr(feat), _lib.ptr(ray_id), _lib.ptr(rb),
                                 _lib.ptr(masks), _lib.ptr(g_sigma), _lib.ptr(g_rgb), _lib.ptr(d_feat),
                                 _lib.ptr(d_params), _lib.ptr(d_rb), 16.0, st))
    _lib.check(L.gf_mlp_ray_bias_backward(R, H, _lib.ptr(p), _lib.ptr(dirs), _lib.ptr(emb), _lib.ptr(d_rb),
                                          _lib.ptr(d_params), _lib.ptr(d_emb), st))
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    for name, got, want in (("d_feat", d_feat.float().cpu().numpy() / 128.0, g["d_feat"]),
                            ("d_params", d_params.cpu().numpy(), g["d_params"]), ("d_emb", d_emb.cpu().numpy(), g["d_emb"])):
        print(name, "rel L2 vs the reference MLPNetwork's autograd:", rel(got, want))
        assert rel(got, want) < 3e-3, name        # north star: 1e-2


def test_charbonnier_and_adam_match_reference():
    from gfnerf_b200 import _lib
    L, st = _lib.lib(), _lib.cur_stream()
    g = load("ref_loss_adam")
    R = g["pred"].shape[0]
    g_rgb, loss = torch.empty((R, 3), device="cuda"), torch.zeros(1, device="cuda")
    pred, target = T(g["pred"]), T(g["target"])
    _lib.check(L.gf_charbonnier(R, _lib.ptr(pred), _lib.ptr(target), 1e-6, _lib.ptr(g_rgb),
                                _lib.ptr(loss), st))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * float(g["loss"])
    close(g_rgb, g["g_pred"], 1e-5, "dL/drgb")
    w = T(g["adam_w0"].copy())
    m1, m2 = torch.zeros_like(w), torch.zeros_like(w)
    shadow = torch.empty(w.numel(), dtype=torch.float16, device="cuda")
    for k, gr in enumerate(g["adam_grads"]):
        grad = T(gr.copy())
        _lib.check(L.gf_adam_step(w.numel(), _lib.ptr(w), _lib.ptr(grad), _lib.ptr(m1), _lib.ptr(m2), _lib.ptr(shadow),
                                  1e-2, 0.9, 0.999, 1e-15, k + 1, 1.0, 1, st))
        close(w, g["adam_traj"][k], 2e-6, f"adam step {k + 1}")
        assert float(grad.abs().max()) == 0.0                       # zero_grad
        assert torch.equal(shadow, w.half())                        # fp16 shadow refreshed


def test_s3im_matches_reference():
    from gfnerf_b200 import _lib
    L, st = _lib.lib(), _lib.cur_stream()
    g = load("ref_s3im")
    R = g["src"].shape[0]
    src, tar, idx = T(g["src"]), T(g["tar"]), T(g["index"])
    g_src, loss = torch.zeros((R, 3), device="cuda"), torch.zeros(1, device="cuda")
    _lib.check(L.gf_s3im(R, idx.numel(), _lib.ptr(idx), _lib.ptr(src), _lib.ptr(tar), 32, 4, 4, 1.0, _lib.ptr(g_src),
                         _lib.ptr(loss), st))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * float(g["loss"])
    close(g_src, g["g_src"], 1e-4, "dL/dsrc")        # fp32 variance terms cancel: 1e-4 of the largest gradient
    # accumulates on top of the Charbonnier gradient / loss, with the multiplier applied
    pred, target = T(g["src"]), T(g["tar"])
    g2, l2 = torch.empty((R, 3), device="cuda"), torch.zeros(1, device="cuda")
    _lib.check(L.gf_charbonnier(R, _lib.ptr(pred), _lib.ptr(target), 1e-6, _lib.ptr(g2), _lib.ptr(l2), st))
    base_g, base_l = g2.clone(), float(l2)
    _lib.check(L.gf_s3im(R, idx.numel(), _lib.ptr(idx), _lib.ptr(pred), _lib.ptr(target), 32, 4, 4, 0.5, _lib.ptr(g2),
                         _lib.ptr(l2), st))
    assert abs(float(l2) - (base_l + 0.5 * float(g["loss"]))) < 1e-5
    close(g2 - base_g, 0.5 * g["g_src"], 1e-4, "scaled dL/dsrc")
