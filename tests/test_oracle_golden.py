"""The oracle against golden vectors produced by RUNNING the reference's own Python code
(tests/golden/make_golden.py: get_weights_f2nerf, the renderers, trunc_exp, MLPNetwork, CharbonnierLoss,
torch.optim.Adam).  CPU only.  This is what pins the composite / MLP / loss / optimizer part of the oracle;
Hash3DAnchored and PersSampler are pinned separately, by the reference's own device code (tests/test_ref_kernels.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5   # BASELINE.json north_star: 1e-5 relative for fp32 values


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: d[k] for k in d.files}


def close(a, b, rtol=RTOL, atol=0.0, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(float(np.abs(b).max()), 1e-30)
    err = float(np.abs(a - b).max())
    assert err <= rtol * scale + atol, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


def to_csr(g):
    counts = g["counts"]
    S = g["sigma"].shape[1]
    m = np.arange(S)[None, :] < counts[:, None]
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return m, offsets


def test_composite_forward_matches_reference():
    g = load("ref_composite")
    m, offsets = to_csr(g)
    out = orc.composite_forward(offsets, g["sigma"][m], g["delta"][m], g["rgb"][m], g["t"][m])
    close(out["weights"], g["weights"][m], what="weights")
    close(out["alphas"], g["alphas"][m], what="alphas")
    close(out["trans"], g["trans"][m], what="transmittance")
    close(out["rgb"], g["out_rgb"], what="rgb")
    close(out["acc"], g["out_acc"], what="accumulation")
    # the reference pads with zero weight, so its padded slots must carry none of the sums
    assert np.all(g["weights"][~m] == 0)
    # DepthRenderer('expected') clips to the global [min t, max t] of the DENSE tensor (renderers.py:281); min is the
    # padding's 0
    depth = np.clip(out["depth"], 0.0, g["t"].max())
    close(depth, g["out_depth"], what="depth")
    # eval-mode RGBRenderer: nan_to_num + clamp (renderers.py:131-137)
    close(np.clip(np.nan_to_num(out["rgb"]), 0, 1), g["out_rgb_eval"], what="rgb eval")


def test_composite_backward_matches_reference_autograd():
    g = load("ref_composite")
    m, offsets = to_csr(g)
    d_sigma, d_rgb = orc.composite_backward(offsets, g["sigma"][m], g["delta"][m], g["rgb"][m], g["g_rgb"], g["g_acc"])
    close(d_sigma, g["d_sigma"][m], rtol=2e-5, what="d_sigma")
    close(d_rgb, g["d_rgb"][m], what="d_rgb")


def test_mlp_forward_backward_match_reference():
    g = load("ref_mlp")
    H = int(g["H"])
    assert orc.mlp_param_count(H) == g["params"].size
    np.testing.assert_array_equal(orc.sh4(g["dirs"]), g["sh"])
    sigma, rgb = orc.mlp_forward(g["params"], g["feat"], g["ray_id"], g["dirs"], g["emb"], H)
    close(sigma, g["sigma"], what="density")
    close(rgb, g["rgb"], what="rgb")
    d_feat, d_params, d_emb = orc.mlp_backward(g["params"], g["feat"], g["ray_id"], g["dirs"], g["emb"], g["g_sigma"],
                                               g["g_rgb"], H)
    close(d_feat, g["d_feat"], rtol=2e-5, what="d_feat")
    # the SH columns of W2 get a gradient in the reference too (SH is an input there, a constant here): all columns
    close(d_params, g["d_params"], rtol=5e-5, what="d_params")
    close(d_emb, g["d_emb"], rtol=2e-5, what="d_emb")


def test_charbonnier_and_adam_match_reference():
    g = load("ref_loss_adam")
    loss, grad = orc.charbonnier(g["pred"], g["target"], 1e-6)
    assert abs(loss - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    close(grad, g["g_pred"], what="dL/drgb")
    w = g["adam_w0"].copy()
    m1, m2 = np.zeros_like(w), np.zeros_like(w)
    for k, gr in enumerate(g["adam_grads"]):
        orc.adam_step(w, gr, m1, m2, 1e-2, 0.9, 0.999, 1e-15, k + 1)
        close(w, g["adam_traj"][k], rtol=2e-6, what=f"adam step {k + 1}")


def test_s3im_matches_reference():
    g = load("ref_s3im")
    loss, grad = orc.s3im(g["src"], g["tar"], g["index"], patch_h=32, ksize=4, stride=4, mult=1.0)
    assert abs(loss - float(g["loss"])) <= 1e-6 * float(g["loss"])
    close(grad, g["g_src"], what="dL/dsrc")
    # the multiplier scales loss and gradient; no gradient requested -> same loss
    l2, g2 = orc.s3im(g["src"], g["tar"], g["index"], mult=0.25)
    assert abs(l2 - 0.25 * loss) < 1e-12 and np.allclose(g2, 0.25 * grad, rtol=1e-6, atol=1e-12)
    assert orc.s3im(g["src"], g["tar"], g["index"], want_grad=False)[0] == loss


def test_field_wiring_matches_the_reference_field_class():
    """The oracle chain hash -> MLPs against the reference's own GFNeRFField.get_density / get_outputs
    (gfnerf/nerfacto_field.py:412-591) run on the CPU with stand-ins only for its native extension (whose AnchoredQuery
    ran the reference's own forward kernel on the host) and for tinycudann (tests/golden/make_golden_field.py)."""
    g = load("ref_field")
    R, S = g["warp_pts"].shape[:2]
    n_vol = g["prim"].shape[1]
    # nerfacto_field.py:431,437: the field queries EVERY slot -- the padding anchor 0 passes `anchors > -1` -- at
    # (warp + 1.5) / 3 with anchor column 0
    assert g["query_pts"].shape[0] == R * S
    assert np.array_equal(g["query_anchors"], g["anchors"][..., 0].reshape(-1))
    pts01 = ((g["warp_pts"].reshape(-1, 3) + np.float32(1.5)) * (np.float32(1.0) / np.float32(3.0))).astype(np.float32)
    # torch's CPU true division vs the multiply-by-reciprocal of its CUDA kernels (what the GPU path reproduces)
    assert np.abs(pts01 - g["query_pts"]).max() <= 2 ** -23
    bias = np.zeros((16 * n_vol, 3), np.float32)
    feat = orc.hash_forward(g["table"], g["prim"], bias, g["query_pts"], g["query_anchors"])
    assert np.array_equal(feat, g["hash_feats"].astype(np.float32))             # the reference's forward kernel
    ray_id = np.repeat(np.arange(R), S).astype(np.int32)
    ray_emb = g["emb"][g["cam"]]                                                # Embedding(rel_camera_indices), :529-531
    sigma, rgb = orc.mlp_forward(g["params"], feat, ray_id, g["dirs"], ray_emb, 64)
    # padding is not masked out by the field (the mask is all true): density = trunc_exp(h0 + 1) everywhere, :499-505
    close(sigma, g["density"].reshape(-1), what="density")
    close(rgb, g["rgb"].reshape(-1, 3), what="rgb")
    d_feat, d_params, d_emb = orc.mlp_backward(g["params"], feat, ray_id, g["dirs"], ray_emb, g["g_sigma"].reshape(-1),
                                               g["g_rgb"].reshape(-1, 3), 64)
    close(d_params, g["d_params"], rtol=5e-5, what="d_params")
    d_emb_cam = np.zeros_like(g["d_emb"], dtype=np.float64)
    np.add.at(d_emb_cam, g["cam"], d_emb)
    close(d_emb_cam, g["d_emb"], rtol=5e-5, what="d_embedding")
