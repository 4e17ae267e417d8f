"""GPU parity: gf_composite_* against the CPU oracle (fp64 prefix sums) and against torch autograd of the
reference formula (nerfstudio/cameras/rays.py:188-200 + renderers.py), ragged CSR incl. empty rays.
Tolerance: 1e-5 relative (north star, fp32 values)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def ragged(R, max_len, seed):
    rng = np.random.RandomState(seed)
    counts = rng.randint(0, max_len + 1, size=R)
    counts[rng.rand(R) < 0.1] = 0
    counts[0] = max_len
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    V = int(offsets[-1])
    sigma = np.exp(rng.normal(0, 2, size=V)).astype(np.float32)
    delta = rng.uniform(1e-3, 5e-2, size=V).astype(np.float32)
    rgb = rng.rand(V, 3).astype(np.float32)
    t = np.sort(rng.uniform(0.1, 30, size=V)).astype(np.float32)
    return offsets, sigma, delta, rgb, t


def run_fwd(offsets, sigma, delta, rgb, t):
    from gfnerf_b200 import _lib
    R, V = offsets.shape[0] - 1, sigma.shape[0]
    dev = "cuda"
    T = lambda a: torch.from_numpy(a).to(dev)
    o, s, d, c, tt = T(offsets), T(sigma), T(delta), T(rgb), T(t)
    w, a, tr = (torch.empty(V, device=dev) for _ in range(3))
    out_rgb, depth, acc, tmax = torch.empty((R, 3), device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev), torch.zeros(1, device=dev)
    _lib.check(_lib.lib().gf_composite_forward(R, _lib.ptr(o), _lib.ptr(s), _lib.ptr(d), _lib.ptr(c), _lib.ptr(tt),
                                               _lib.ptr(w), _lib.ptr(a), _lib.ptr(tr), _lib.ptr(out_rgb),
                                               _lib.ptr(depth), _lib.ptr(acc), _lib.ptr(tmax), _lib.cur_stream()))
    return dict(weights=w, alphas=a, trans=tr, rgb=out_rgb, depth=depth, acc=acc, tmax=tmax), (o, s, d, c, tt)


@pytest.mark.parametrize("R,max_len,seed", [(300, 1024, 0), (4096, 90, 1), (5, 33, 2)])
def test_forward_backward_match_oracle(R, max_len, seed):
    from gfnerf_b200 import _lib
    offsets, sigma, delta, rgb, t = ragged(R, max_len, seed)
    got, (o, s, d, c, tt) = run_fwd(offsets, sigma, delta, rgb, t)
    ref = orc.composite_forward(offsets, sigma, delta, rgb, t)
    for k in ("weights", "alphas", "trans", "rgb", "depth", "acc"):
        np.testing.assert_allclose(got[k].cpu().numpy(), ref[k], rtol=1e-5, atol=1e-6, err_msg=k)
    assert float(got["tmax"]) == t.max()
    rng = np.random.RandomState(seed + 10)
    g_rgb = rng.normal(size=(R, 3)).astype(np.float32)
    g_acc = rng.normal(size=R).astype(np.float32)
    V = sigma.shape[0]
    d_sigma, d_rgb = torch.empty(V, device="cuda"), torch.empty((V, 3), device="cuda")
    tg, ta = torch.from_numpy(g_rgb).cuda(), torch.from_numpy(g_acc).cuda()
    _lib.check(_lib.lib().gf_composite_backward(R, _lib.ptr(o), _lib.ptr(s), _lib.ptr(d), _lib.ptr(c),
                                                _lib.ptr(got["trans"]), _lib.ptr(tg), _lib.ptr(ta), None,
                                                _lib.ptr(d_sigma), _lib.ptr(d_rgb), _lib.cur_stream()))
    rs, rc = orc.composite_backward(offsets, sigma, delta, rgb, g_rgb, g_acc)
    scale = np.abs(rs).max()
    # w = alpha * exp(-prefix): the fp32 prefix sum's absolute error becomes a relative error of w deep in the ray
    np.testing.assert_allclose(d_rgb.cpu().numpy(), rc, rtol=1e-5, atol=1e-6 * np.abs(rc).max())
    # d_sigma is a difference of two sums: 1e-5 relative to the ray's gradient scale
    assert np.all(np.abs(d_sigma.cpu().numpy() - rs) <= 1e-5 * np.abs(rs) + 2e-6 * scale)


def test_against_torch_autograd_of_reference_formula():
    """Dense [R,S,1] get_weights_f2nerf + renderers in float64 torch, incl. the zero-delta padding."""
    from gfnerf_b200 import _lib
    R, S = 64, 128
    rng = np.random.RandomState(7)
    counts = rng.randint(1, S + 1, size=R)
    dense_sigma = np.exp(rng.normal(0, 1.5, size=(R, S, 1)))
    dense_delta = rng.uniform(1e-3, 5e-2, size=(R, S, 1))
    dense_rgb = rng.rand(R, S, 3)
    dense_t = np.sort(rng.uniform(0.1, 10, size=(R, S, 1)), axis=1)
    pad = np.arange(S)[None, :, None] >= counts[:, None, None]
    dense_delta[pad] = 0.0
    dense_t[pad] = 0.0
    sg = torch.tensor(dense_sigma, requires_grad=True)
    cl = torch.tensor(dense_rgb, requires_grad=True)
    dl, tt = torch.tensor(dense_delta), torch.tensor(dense_t)
    dd = dl * sg                                                   # rays.py:188-200
    alphas = 1 - torch.exp(-dd)
    trans = torch.cumsum(dd[..., :-1, :], dim=-2)
    trans = torch.cat([torch.zeros((R, 1, 1), dtype=trans.dtype), trans], dim=-2)
    trans = torch.exp(-trans)
    w = torch.nan_to_num(alphas * trans)
    out_rgb = (w * cl).sum(-2)                                      # renderers.py:97-110
    acc = w.sum(-2)                                                 # :220
    depth = (w * tt).sum(-2) / (acc + 1e-10)                        # :269-280
    g_rgb, g_acc = torch.tensor(rng.normal(size=(R, 3))), torch.tensor(rng.normal(size=(R, 1)))
    ((out_rgb * g_rgb).sum() + (acc * g_acc).sum()).backward()
    m = ~pad[..., 0]
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    f32 = lambda a: a.astype(np.float32)
    got, (o, s, d, c, t_) = run_fwd(offsets, f32(dense_sigma[..., 0][m]), f32(dense_delta[..., 0][m]), f32(dense_rgb[m]),
                                    f32(dense_t[..., 0][m]))
    np.testing.assert_allclose(got["rgb"].cpu().numpy(), out_rgb.detach().numpy(), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(got["acc"].cpu().numpy(), acc.detach().numpy()[:, 0], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(got["depth"].cpu().numpy(), depth.detach().numpy()[:, 0], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(got["weights"].cpu().numpy(), w.detach().numpy()[..., 0][m], rtol=2e-5, atol=1e-7)
    V = int(offsets[-1])
    d_sigma, d_rgb = torch.empty(V, device="cuda"), torch.empty((V, 3), device="cuda")
    tg = torch.from_numpy(f32(g_rgb.numpy())).cuda()
    ta = torch.from_numpy(f32(g_acc.numpy()[:, 0])).cuda()
    _lib.check(_lib.lib().gf_composite_backward(R, _lib.ptr(o), _lib.ptr(s), _lib.ptr(d), _lib.ptr(c),
                                                _lib.ptr(got["trans"]), _lib.ptr(tg), _lib.ptr(ta), None,
                                                _lib.ptr(d_sigma), _lib.ptr(d_rgb), _lib.cur_stream()))
    rs = sg.grad.numpy()[..., 0][m]
    np.testing.assert_allclose(d_rgb.cpu().numpy(), cl.grad.numpy()[m], rtol=2e-5, atol=2e-6 * np.abs(cl.grad.numpy()).max())
    assert np.all(np.abs(d_sigma.cpu().numpy() - rs) <= 2e-5 * np.abs(rs) + 4e-6 * np.abs(rs).max())
    # weights-only call (operator API) and gradient through g_w
    w2 = torch.empty(V, device="cuda")
    tr2 = torch.empty(V, device="cuda")
    _lib.check(_lib.lib().gf_composite_forward(R, _lib.ptr(o), _lib.ptr(s), _lib.ptr(d), None, None, _lib.ptr(w2), None,
                                               _lib.ptr(tr2), None, None, None, None, _lib.cur_stream()))
    assert torch.equal(w2, got["weights"])
    gw = torch.from_numpy(f32(rng.normal(size=V))).cuda()
    ds2 = torch.empty(V, device="cuda")
    _lib.check(_lib.lib().gf_composite_backward(R, _lib.ptr(o), _lib.ptr(s), _lib.ptr(d), None, _lib.ptr(tr2), None, None,
                                                _lib.ptr(gw), _lib.ptr(ds2), None, _lib.cur_stream()))
    sg2 = torch.tensor(dense_sigma, requires_grad=True)
    dd = dl * sg2
    tr = torch.exp(-torch.cat([torch.zeros((R, 1, 1), dtype=dd.dtype), torch.cumsum(dd[..., :-1, :], dim=-2)], dim=-2))
    wd = (1 - torch.exp(-dd)) * tr
    gw_dense = torch.zeros((R, S), dtype=torch.float64)
    gw_dense[torch.from_numpy(m)] = gw.cpu().double()
    (wd[..., 0] * gw_dense).sum().backward()
    rs = sg2.grad.numpy()[..., 0][m]
    assert np.all(np.abs(ds2.cpu().numpy() - rs) <= 2e-5 * np.abs(rs) + 4e-6 * np.abs(rs).max())
