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


def test_whole_path_matches_the_reference_model_code():
    """The oracle chain sampler -> hash -> MLPs -> composite -> octree vote against the reference's own Python model
    path -- PersSampler.generate_ray_samples, GFNeRFField, get_weights_f2nerf, the renderers, GFNeRFModel.get_outputs --
    run on the CPU over the reference's kernels compiled for the host (tests/golden/make_golden_model.py)."""
    from tests.helpers import load_rig
    g = load("ref_model")
    rig = load_rig("rig8")
    R, S = g["rays_o"].shape[0], 1024
    assert bool(g["deltas_are_dists"]) and bool(g["starts_are_t"])        # perssampler.py:411-412, 429
    assert list(g["calls"]) == [f"UpdateRayMarch:{int(g['step'])}", "UpdateMode:0", f"UpdateOctNodes:{int(g['step'])}"]
    # GetSamples normalises the directions (PersSampler_cuda.cu:323)
    d_unit = g["rays_d_raw"] / np.linalg.norm(g["rays_d_raw"], axis=1, keepdims=True)
    assert np.abs(d_unit - g["rays_d_unit"]).max() < 2e-7
    smp = orc.sampler_get_samples(g["rays_o"], g["rays_d_unit"], g["noise"], rig["tree_nodes"], rig["pers_trans"])
    counts = smp["counts"]
    assert np.array_equal(counts, g["counts"]) and counts[-1] == 0
    m = counts[:, None] > np.arange(S)[None]
    pts01 = ((smp["warp_pts"][m] + np.float32(1.5)) * (np.float32(1.0) / np.float32(3.0))).astype(np.float32)
    n_vol = g["prim"].shape[1]
    feat = orc.hash_forward(g["table"], g["prim"], np.zeros((16 * n_vol, 3), np.float32), pts01, smp["anchors"][m][:, 0])
    ray_id = np.repeat(np.arange(R), counts).astype(np.int32)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    ray_emb = g["emb"][g["cam"]]
    # The field encodes the bundle's directions AS GIVEN: generate_ray_samples puts ray_bundle.directions into the
    # frustums (perssampler.py:414-418) and get_outputs feeds (directions + 1) / 2 to the SH (nerfacto_field.py:518-521);
    # only GetSamples normalises, for the march.  The fixture's bundle is deliberately not normalised.
    sigma, rgb = orc.mlp_forward(g["params"], feat, ray_id, g["rays_d_raw"], ray_emb, 64)
    sigma_u, rgb_u = orc.mlp_forward(g["params"], feat, ray_id, g["rays_d_unit"], ray_emb, 64)
    assert np.array_equal(sigma, sigma_u) and np.abs(rgb - rgb_u).max() > 1e-3      # (so the distinction is visible)
    comp = orc.composite_forward(offsets, sigma, smp["dists"][m], rgb, smp["ts"][m])
    # 1e-4: the reference's sample positions come from the host build of its march kernel (g++'s contraction), the
    # oracle's follow nvcc's; see tests/test_ref_kernels.py
    close(comp["rgb"], g["out_rgb"], rtol=1e-4, what="rgb")
    close(comp["acc"], g["out_acc"][:, 0], rtol=1e-4, what="accumulation")
    # DepthRenderer('expected') clips to the dense tensor's [min, max] of t (min = the padding's 0), then / scale_factor
    depth = np.clip(comp["depth"], 0.0, smp["ts"].max()) / g["scale_factor"]
    close(depth, g["out_depth"][:, 0], rtol=1e-4, what="depth")
    close(smp["first_oct_dis"] / g["scale_factor"], g["out_oct_depth"][:, 0], what="oct_depth")
    # backward of mean((rgb - target)^2) through composite and both MLPs
    g_rgb = (2.0 * (comp["rgb"] - g["target"]) / np.float32(R * 3)).astype(np.float32)
    assert abs(float(((comp["rgb"] - g["target"]) ** 2).mean()) - float(g["loss"])) <= 1e-4 * float(g["loss"])
    d_sigma, d_rgb = orc.composite_backward(offsets, sigma, smp["dists"][m], rgb, g_rgb)
    _, d_params, d_emb = orc.mlp_backward(g["params"], feat, ray_id, g["rays_d_raw"], ray_emb, d_sigma, d_rgb, 64)
    close(d_params, g["d_params"], rtol=2e-3, what="d_params")
    d_emb_cam = np.zeros_like(g["d_emb"], dtype=np.float64)
    np.add.at(d_emb_cam, g["cam"], d_emb)
    close(d_emb_cam, g["d_emb"], rtol=2e-3, what="d_embedding")
    # the octree vote with the arguments get_outputs passes (nerfacto.py:603-611)
    w_dense, a_dense = np.zeros((R, S), np.float32), np.zeros((R, S), np.float32)
    w_dense[m], a_dense[m] = comp["weights"], comp["alphas"]
    nodes = rig["tree_nodes"].copy()
    n = nodes.size // 128
    ws, as_, vc = np.full(n, 1000, np.int64), np.full(n, 1000, np.int64), np.zeros(n, np.int64)
    orc.update_oct_nodes(counts, smp["anchors"][..., 1].reshape(-1), w_dense.reshape(-1), a_dense.reshape(-1), nodes, ws,
                         as_, vc)
    assert np.array_equal(vc, g["visit"])
    # a vote flips when a leaf's max weight sits within 1e-4 of its threshold: allow none here, report if it happens
    assert np.array_equal(ws, g["w_stats"]) and np.array_equal(as_, g["a_stats"])
    assert np.array_equal(nodes.view(np.int64).reshape(-1, 16)[:, 12], g["trans_idx"])
