"""GPU parity: fused field MLP (gf_mlp_*) against the fp32 CPU oracle of the reference field
(gfnerf/nerfacto_field.py:437-591 with gfnerf/mlp.py).  Tolerance: 1e-2 relative -- the "fp16 MLP" class of the
north star (fp16 weights / activations, fp32 accumulate)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
WIDTHS = [64, 128]     # nerfstudio's default width / the reference's shipped gf-nerf config (gfnerf/config.py:124-125)


def offsets(H):
    """name -> (start, end) in the parameter blob (include/gfnerf_b200.h)"""
    out, o = {}, 0
    for name, (a, b) in zip(("0", "1", "2", "3", "4"), ((H, 32), (16, H), (H, 63), (H, H), (3, H))):
        out["w" + name] = (o, o + a * b); o += a * b
        out["b" + name] = (o, o + a); o += a
    out["count"] = o
    return out


def make_case(n, R, seed, with_emb=True, ray_stride=1, H=64):
    """ray_stride > 1 leaves ray ids without samples in between (rays that missed the octree)"""
    rng = np.random.RandomState(seed)
    g = torch.Generator().manual_seed(seed)
    from gfnerf_b200.engine import init_mlp_params
    params = init_mlp_params(H, g).numpy()
    # torch-default init gives tiny outputs for 0.01-scale hash features; scale features up so that every layer is
    # exercised in a non-trivial range (post-training features are O(0.1 - 1))
    feat = (rng.normal(0, 0.5, size=(n, 32))).astype(np.float16)
    cuts = np.sort(rng.choice(np.arange(1, n), size=R - 1, replace=False)) if R > 1 else np.array([], int)
    ray_id = (np.searchsorted(cuts, np.arange(n), side="right") * ray_stride).astype(np.int32)
    R = (R - 1) * ray_stride + 1
    dirs = rng.normal(size=(R, 3))
    dirs = (dirs / np.linalg.norm(dirs, axis=-1, keepdims=True)).astype(np.float32)
    emb = rng.normal(size=(R, 32)).astype(np.float32) if with_emb else None
    return params, feat, ray_id, dirs, emb


def _split_params(p, H):
    o, out = 0, []
    for a, b in ((H, 32), (16, H), (H, 63), (H, H), (3, H)):
        out.append(p[o:o + a * b].reshape(a, b)); o += a * b
        out.append(p[o:o + a]); o += a
    return out


Q = lambda x: x.astype(np.float16).astype(np.float32)


def Qhl(x):
    """value carried as an fp16 pair hi + lo, hi rounded to nearest (the weight / bias tiles)"""
    hi = Q(x)
    return hi + Q(x - hi)


def Qtr(x):
    """value carried as an fp16 pair hi + lo, hi TRUNCATED to 11 significant bits (the activations, csrc/mlp_tc.cu
    trunc11)"""
    hi = (np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    return Q(hi) + Q(x - hi)


def emulate_split_forward(params, feat, ray_id, dirs, emb, H, bias_fp32=False):
    """The kernel's split-precision forward in numpy: weights, biases and hidden activations of layers 0..3 as fp16
    pairs, fp32 accumulate; the output layer takes h3 as plain fp16.  Returns (sigma, rgb, relu masks)."""
    w0, b0, w1, b1, w2, b2, w3, b3, w4, b4 = _split_params(params, H)
    Qb = (lambda b: b) if bias_fp32 else Qhl      # H = 128 adds fp32 biases in the epilogue, H = 64 a bias MMA (hi + lo)
    x = feat.astype(np.float32)
    p1 = x @ Qhl(w0).T + Qb(b0)
    h1 = Qtr(np.maximum(p1, 0))
    h = Qtr(h1 @ Qhl(w1).T + Qb(b1))
    rb = orc.sh4(dirs) @ w2[:, :16].T + b2 + (emb @ w2[:, 31:].T if emb is not None else 0)
    p2 = h[:, 1:] @ Qhl(w2[:, 16:31]).T + rb[ray_id]
    h2 = Qtr(np.maximum(p2, 0))
    p3 = h2 @ Qhl(w3).T + Qb(b3)
    h3 = Q(np.maximum(p3, 0))
    o = h3 @ Q(w4).T + Qb(b4)
    return np.exp(h[:, 0] + 1), 1 / (1 + np.exp(-o)), (p1 > 0, p2 > 0, p3 > 0)


def emulate_fp16_backward(params, feat, ray_id, dirs, emb, d_sigma, d_rgb, H):
    """numpy restatement of the reference field's backward (orc.mlp_backward) with fp16 rounding at the points where
    the kernel holds fp16 (weights, activations and gradient fragments; biases as fp16 pairs); fp32 accumulation.
    The ReLU masks are the split-precision forward's, as in the kernel."""
    w0, b0, w1, b1, w2, b2, w3, b3, w4, b4 = _split_params(params, H)
    _, _, (m1, m2, m3) = emulate_split_forward(params, feat, ray_id, dirs, emb, H, bias_fp32=H == 128)
    # b0, b1, b3, b4 enter through a bias MMA (two fp16 columns: hi + lo); b2 is part of the fp32 per-ray bias
    if H == 64:
        b0, b1, b3, b4 = Qhl(b0), Qhl(b1), Qhl(b3), Qhl(b4)
    x = feat.astype(np.float32)
    h1 = Q(np.maximum(x @ Q(w0).T + b0, 0))
    h = h1 @ Q(w1).T + b1
    rb = orc.sh4(dirs) @ w2[:, :16].T + b2 + (emb @ w2[:, 31:].T if emb is not None else 0)
    h2 = Q(np.maximum(Q(h)[:, 1:] @ Q(w2[:, 16:31]).T + rb[ray_id], 0))
    h3 = Q(np.maximum(h2 @ Q(w3).T + b3, 0))
    o = h3 @ Q(w4).T + b4
    sg = 1 / (1 + np.exp(-o))
    gh3 = (Q(d_rgb * sg * (1 - sg)) @ Q(w4)) * m3
    gh2 = (Q(gh3) @ Q(w3)) * m2
    gh = np.concatenate([(d_sigma * np.exp(np.clip(h[:, 0] + 1, -15, 15)))[:, None], Q(gh2) @ Q(w2[:, 16:31])], 1)
    gh1 = (Q(gh) @ Q(w1)) * m1
    go = Q(d_rgb * sg * (1 - sg))
    # weight gradients from the fp16 gradient / activation tiles (fp32 accumulate); W2's SH / emb columns and b2 come
    # from the fp32 per-ray sums of g h2
    R = dirs.shape[0]
    G = np.zeros((R, H), np.float64)
    np.add.at(G, ray_id, gh2.astype(np.float64))
    per_ray_in = np.concatenate([orc.sh4(dirs), np.zeros((R, 15), np.float32), emb if emb is not None else np.zeros((R, 32), np.float32)], 1)
    dw2 = G.T @ per_ray_in
    dw2[:, 16:31] = Q(gh2).T.astype(np.float64) @ Q(h)[:, 1:]
    dp = np.concatenate([
        (Q(gh1).T.astype(np.float64) @ x).reshape(-1), Q(gh1).sum(0, dtype=np.float64),
        (Q(gh).T.astype(np.float64) @ h1).reshape(-1), Q(gh).sum(0, dtype=np.float64),
        dw2.reshape(-1), G.sum(0),
        (Q(gh3).T.astype(np.float64) @ h2).reshape(-1), Q(gh3).sum(0, dtype=np.float64),
        (go.T.astype(np.float64) @ h3).reshape(-1), go.sum(0, dtype=np.float64)])
    demb = G @ w2[:, 31:].astype(np.float64) if emb is not None else None
    return Q(gh1) @ Q(w0), dp, demb


def gpu_forward(params, feat, ray_id, dirs, emb, H):
    from gfnerf_b200 import _lib
    L, st = _lib.lib(), _lib.cur_stream()
    n, R = feat.shape[0], dirs.shape[0]
    T = lambda a: None if a is None else torch.from_numpy(a).cuda()
    tp, tf, tr, td, te = T(params), T(feat), T(ray_id), T(dirs), T(emb)
    rb = torch.empty((R, H), device="cuda")
    _lib.check(L.gf_mlp_ray_bias(R, H, _lib.ptr(tp), _lib.ptr(td), _lib.ptr(te), _lib.ptr(rb), st))
    sigma, rgb = torch.empty(n, device="cuda"), torch.empty((n, 3), device="cuda")
    masks = torch.zeros((n, 2, H // 16), dtype=torch.int32, device="cuda")
    assert L.gf_mlp_mask_words(H) == 2 * (H // 16)
    _lib.check(L.gf_mlp_forward(n, None, H, _lib.ptr(tp), _lib.ptr(tf), _lib.ptr(tr), _lib.ptr(rb), _lib.ptr(sigma),
                                _lib.ptr(rgb), _lib.ptr(masks), st))
    return (tp, tf, tr, td, te, rb, masks), sigma, rgb


def unpack_masks(masks, layer, H):
    """relu_masks uint32 [n][2][H / 16] -> bool [n, H] of hidden layer `layer`: per column half, one word per 32 columns
    (H = 64: word `layer`; H = 128: words 2 layer, 2 layer + 1), bits as csrc/mlp_tc_common.cuh mask_bits_of_pair puts
    them: the two bits of column pair q sit at 7 - q and 23 - q for q < 8, at 15 - (q - 8) and 31 - (q - 8) above"""
    w = masks.cpu().numpy().view(np.uint32)
    half, chunks = H // 2, H // 64
    out = np.zeros((w.shape[0], H), bool)
    for hf in range(2):
        for c in range(chunks):
            word = w[:, hf, chunks * layer + c]
            for q in range(16):
                b0, b1 = (7 - q, 23 - q) if q < 8 else (15 - (q - 8), 31 - (q - 8))
                out[:, half * hf + 32 * c + 2 * q] = (word >> np.uint32(b0)) & 1
                out[:, half * hf + 32 * c + 2 * q + 1] = (word >> np.uint32(b1)) & 1
    return out


def test_param_count_and_unsupported_width():
    from gfnerf_b200 import _lib
    assert _lib.lib().gf_mlp_param_count(64) == orc.mlp_param_count(64) == 11603 == offsets(64)["count"]
    assert _lib.lib().gf_mlp_param_count(128) == orc.mlp_param_count(128) == 31379 == offsets(128)["count"]
    assert _lib.lib().gf_mlp_param_count(48) == -1 and _lib.lib().gf_mlp_mask_words(48) == -1


@pytest.mark.parametrize("H", WIDTHS)
@pytest.mark.parametrize("n,R,seed,with_emb", [(5000, 37, 0, True), (31, 2, 1, False), (4096, 1, 2, True)])
def test_forward_matches_oracle(n, R, seed, with_emb, H):
    params, feat, ray_id, dirs, emb = make_case(n, R, seed, with_emb, H=H)
    (tp, tf, tr, td, te, rb, masks), sigma, rgb = gpu_forward(params, feat, ray_id, dirs, emb, H)
    # per-ray bias: fp32 against the oracle's SH and the reference weight slices
    sh = orc.sh4(dirs)
    off = offsets(H)
    w2 = params[slice(*off["w2"])].reshape(H, 63)
    b2 = params[slice(*off["b2"])]
    ref_rb = sh.astype(np.float64) @ w2[:, :16].T.astype(np.float64) + b2
    if emb is not None:
        ref_rb = ref_rb + emb.astype(np.float64) @ w2[:, 31:].T.astype(np.float64)
    np.testing.assert_allclose(rb.cpu().numpy(), ref_rb, rtol=1e-5, atol=1e-5)
    ref_sigma, ref_rgb = orc.mlp_forward(params, feat.astype(np.float32), ray_id, dirs, emb, H)
    # split-precision forward: hidden layers to ~1e-6, the fp16 output layer to ~1e-3 -- far inside the 1e-2 of the
    # north star's "fp16 MLP" clause
    np.testing.assert_allclose(sigma.cpu().numpy(), ref_sigma, rtol=2e-4)
    np.testing.assert_allclose(rgb.cpu().numpy(), ref_rgb, rtol=2e-3, atol=2e-4)
    print("max rel err sigma", np.max(np.abs(sigma.cpu().numpy() - ref_sigma) / ref_sigma),
          "max abs err rgb", np.max(np.abs(rgb.cpu().numpy() - ref_rgb)))
    # the ReLU masks handed to the backward: those of the fp32 reference except where a pre-activation is within
    # ~1e-5 of zero, and exactly those of the numpy restatement of the split arithmetic away from such ties
    emu_sigma, emu_rgb, emu_masks = emulate_split_forward(params, feat, ray_id, dirs, emb, H, bias_fp32=H == 128)
    np.testing.assert_allclose(sigma.cpu().numpy(), emu_sigma, rtol=1e-5)
    np.testing.assert_allclose(rgb.cpu().numpy(), emu_rgb, rtol=1e-4, atol=1e-5)
    w0, b0, w1, b1, w2, b2, w3, b3, w4, b4 = _split_params(params, H)
    x = feat.astype(np.float64)
    p1 = x @ w0.T.astype(np.float64) + b0
    h = np.maximum(p1, 0) @ w1.T.astype(np.float64) + b1
    p2 = h[:, 1:] @ w2[:, 16:31].T.astype(np.float64) + ref_rb[ray_id]
    p3 = np.maximum(p2, 0) @ w3.T.astype(np.float64) + b3
    for layer, pre in enumerate((p1, p2, p3)):
        got = unpack_masks(masks, layer, H)
        flips = got != (pre > 0)
        print(f"layer {layer}: {flips.sum()} of {flips.size} ReLU masks differ from the fp64 reference; "
              f"largest |pre-activation| among them {np.abs(pre[flips]).max() if flips.any() else 0:.2e}")
        assert flips.mean() < 2e-5
        assert not flips.any() or np.abs(pre[flips]).max() < 2e-5 * max(1.0, np.abs(pre).max())


@pytest.mark.parametrize("H", WIDTHS)
def test_device_side_count_limits_work(H):
    from gfnerf_b200 import _lib
    params, feat, ray_id, dirs, emb = make_case(1000, 5, 3, H=H)
    (tp, tf, tr, td, te, rb, masks), sigma, rgb = gpu_forward(params, feat, ray_id, dirs, emb, H)
    n_dev = torch.tensor([613], dtype=torch.int32, device="cuda")
    for split in (True, False):          # training forward (split precision + masks) / inference forward (plain fp16)
        s2, c2 = torch.zeros(1000, device="cuda"), torch.zeros((1000, 3), device="cuda")
        m2 = torch.zeros((1000, 2, H // 16), dtype=torch.int32, device="cuda") if split else None
        _lib.check(_lib.lib().gf_mlp_forward(1000, _lib.ptr(n_dev), H, _lib.ptr(tp), _lib.ptr(tf), _lib.ptr(tr),
                                             _lib.ptr(rb), _lib.ptr(s2), _lib.ptr(c2), _lib.ptr(m2), _lib.cur_stream()))
        if split:
            assert torch.equal(s2[:613], sigma[:613]) and torch.equal(c2[:613], rgb[:613])
            assert torch.equal(m2[:613], masks[:613]) and not m2[613:].any()
        else:   # the two forwards agree to fp16-MLP accuracy
            assert torch.allclose(s2[:613], sigma[:613], rtol=5e-3) and torch.allclose(c2[:613], rgb[:613], atol=2e-3)
        assert not s2[613:].any() and not c2[613:].any()


# the per-ray gradient of the head's first-layer bias leaves a 128-sample tile through 8 (H = 128: 16) "ray slots" (an
# MMA against a one-hot matrix) when the tile's ray ids span fewer, else through a CUDA-core path: long rays, rays with
# gaps in their ids (stride 2: slots 0, 2, 4, 6; stride 3: mixed), many short rays (every tile overflows the slots)
@pytest.mark.parametrize("H", WIDTHS)
@pytest.mark.parametrize("n,R,seed,with_emb,stride", [(6000, 41, 5, True, 1), (130, 3, 6, False, 1), (129, 1, 7, True, 1),
                                                      (6000, 41, 5, True, 2), (6000, 60, 5, True, 3),
                                                      (6000, 800, 5, True, 1)])
def test_backward_matches_oracle(n, R, seed, with_emb, stride, H):
    from gfnerf_b200 import _lib
    L, st = _lib.lib(), _lib.cur_stream()
    params, feat, ray_id, dirs, emb = make_case(n, R, seed, with_emb, stride, H=H)
    R = dirs.shape[0]
    off = offsets(H)
    (tp, tf, tr, td, te, rb, masks), sigma, rgb = gpu_forward(params, feat, ray_id, dirs, emb, H)
    rng = np.random.RandomState(seed + 100)
    d_sigma = (rng.normal(size=n) * 1e-4).astype(np.float32)
    d_rgb = (rng.normal(size=(n, 3)) * 1e-4).astype(np.float32)       # the magnitude a mean over ~8k rays produces
    ref_dfeat, ref_dparams, ref_demb = orc.mlp_backward(params, feat.astype(np.float32), ray_id, dirs, emb, d_sigma,
                                                        d_rgb, H)
    tds, tdc = torch.from_numpy(d_sigma).cuda(), torch.from_numpy(d_rgb).cuda()
    d_feat = torch.zeros((n, 32), dtype=torch.float16, device="cuda")
    d_params = torch.zeros(off["count"], device="cuda")
    d_rb = torch.zeros((R, H), device="cuda")
    _lib.check(L.gf_mlp_backward(n, None, H, _lib.ptr(tp), _lib.ptr(tf), _lib.ptr(tr), _lib.ptr(rb), _lib.ptr(masks),
                                 _lib.ptr(tds), _lib.ptr(tdc), _lib.ptr(d_feat), _lib.ptr(d_params), _lib.ptr(d_rb),
                                 8192.0, st))
    d_emb = torch.zeros((R, 32), device="cuda") if with_emb else None
    _lib.check(L.gf_mlp_ray_bias_backward(R, H, _lib.ptr(tp), _lib.ptr(td), _lib.ptr(te), _lib.ptr(d_rb),
                                          _lib.ptr(d_params), _lib.ptr(d_emb), st))
    got_dfeat = d_feat.float().cpu().numpy() / 128.0
    # (1) against the fp32 reference.  The ReLU masks are the split-precision forward's (an fp16 forward flips the
    # mask of ~1e-3 of the hidden units -- those within fp16 rounding of zero -- and every flip is a 100 % error of
    # one gradient term: 1-2 % in relative L2, measured; with the masks right what is left is fp16 rounding of the
    # gradient fragments, ~4e-4)
    s = np.abs(ref_dfeat).max()
    diff = np.abs(got_dfeat - ref_dfeat)
    rel_l2 = np.linalg.norm(diff) / np.linalg.norm(ref_dfeat)
    print(f"d_feat vs fp32 oracle: rel L2 {rel_l2:.2e}, max {diff.max() / s:.2e} of max")
    if n >= 1000:
        assert rel_l2 < 2e-3
    # (2) against the same reference with the kernel's fp16 quantisation points emulated in numpy: tight
    emu, emu_dp, emu_demb = emulate_fp16_backward(params, feat, ray_id, dirs, emb, d_sigma * 8192.0, d_rgb * 8192.0, H)
    emu, emu_dp = emu / 8192.0, emu_dp / 8192.0
    d2 = np.abs(got_dfeat - emu)
    print(f"d_feat vs fp16-emulating oracle: max {d2.max() / s:.2e} of max")
    assert np.all(d2 <= 2e-3 * np.abs(emu) + 2e-3 * s + 2.0 ** -24 / 128)
    got = d_params.cpu().numpy().astype(np.float64)
    names = [(k,) + off[k] for k in ("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4")]
    for name, a, b in names:
        r, e, g = ref_dparams[a:b], emu_dp[a:b], got[a:b]
        sc = np.abs(r).max() + 1e-30
        err_ref, err_emu = np.max(np.abs(g - r)) / sc, np.max(np.abs(g - e)) / sc
        l2 = np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
        print(f"d_{name}: vs fp32 oracle max {err_ref:.2e} / rel L2 {l2:.2e}; vs fp16-emulating oracle max {err_emu:.2e}")
        assert err_emu < 2e-3, name           # the kernel computes exactly the fp16-MLP gradient
        if n >= 1000:
            assert l2 < 2e-3, name            # north star: 1e-2 for the fp16 MLP; measured ~4e-4
    if with_emb:
        sc = np.abs(ref_demb).max()
        got_e = d_emb.cpu().numpy()
        assert np.max(np.abs(got_e - emu_demb / 8192.0)) / sc < 2e-3
        if n >= 1000:
            assert np.linalg.norm(got_e - ref_demb) / np.linalg.norm(ref_demb) < 2e-3
    else:
        w2g = got[slice(*off["w2"])].reshape(H, 63)
        assert not w2g[:, 31:].any()        # no embedding: those columns get no gradient
    # frozen-MLP variant (focal stage): same d_feat, no parameter gradients
    d_feat2 = torch.zeros_like(d_feat)
    _lib.check(L.gf_mlp_backward(n, None, H, _lib.ptr(tp), _lib.ptr(tf), _lib.ptr(tr), _lib.ptr(rb), _lib.ptr(masks),
                                 _lib.ptr(tds), _lib.ptr(tdc), _lib.ptr(d_feat2), None, None, 8192.0, st))
    assert torch.equal(d_feat, d_feat2)
