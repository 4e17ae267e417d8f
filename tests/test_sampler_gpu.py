"""GPU parity: gf_sampler_* (through the C-ABI / PersSamplerCore) against the CPU oracle on the prebuilt rig.

Bar (north star): sample counts, octree node ids and transform ids bit-exact; positions, distances and t within
1e-5 relative (they are expected to be bit-equal too: both sides follow one FMA convention -- reported)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.helpers import load_rig, make_sampler

pytestmark = pytest.mark.gpu


def rays_for(rig, n, seed):
    from gfnerf_b200.persoctree import rig_rays
    return rig_rays(rig["c2w"], rig["intri"], n, seed=seed)


def valid_mask(counts):
    return counts[:, None] > np.arange(1024)[None]


@pytest.mark.parametrize("mode,n_rays", [(1, 777), (0, 2048)])
def test_get_samples_matches_oracle(mode, n_rays):
    rig = load_rig("rig8")
    s = make_sampler(rig, mode=mode)
    o, d, _ = rays_for(rig, n_rays, seed=5 + mode)
    d_raw = d * np.linspace(0.5, 2.0, n_rays, dtype=np.float32)[:, None]      # GetSamples normalises (:323)
    rng = np.random.RandomState(3)
    noise = np.ones(1024 + n_rays + 10, np.float32) if mode == 1 else rng.uniform(0.5, 1.5, 1024 + n_rays + 10).astype(np.float32)
    fineness = 1.0 if mode == 1 else 4.0
    s.ray_march_fineness_ = fineness
    tn = (torch.from_numpy(noise).cuda() * np.float32(fineness)).contiguous()
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d_raw).cuda()
    world, warp, dirs, dists, ts, anchors, start_end, first = s.GetSamples(to, td, None, noise=tn)
    d_unit = (td / torch.linalg.norm(td, 2, -1, True)).cpu().numpy()
    ref = orc.sampler_get_samples(o, d_unit, tn.cpu().numpy(), rig["tree_nodes"], rig["pers_trans"])
    counts = (start_end[:, 1] - start_end[:, 0]).cpu().numpy()
    assert np.array_equal(counts, ref["counts"])                               # bit-exact sample counts
    assert counts.max() > 100 and counts.min() >= 0
    excl = np.concatenate([[0], np.cumsum(ref["counts"])[:-1]])
    assert np.array_equal(start_end[:, 0].cpu().numpy(), excl)
    m = valid_mask(counts)
    assert np.array_equal(anchors.cpu().numpy()[m], ref["anchors"][m])         # bit-exact trans / node / block ids
    for name, got in (("world_pts", world), ("warp_pts", warp), ("dirs", dirs), ("dists", dists), ("ts", ts)):
        g, r = got.cpu().numpy(), ref[name]
        np.testing.assert_allclose(g[m], r[m], rtol=1e-5, atol=1e-6, err_msg=name)
        print(f"{name}: bit-equal fraction {np.mean(g[m] == r[m]):.6f}")
        assert not g[~m].any(), f"{name}: padding must stay zero"
    np.testing.assert_allclose(first.cpu().numpy().reshape(-1), ref["first_oct_dis"], rtol=1e-6)
    if mode == 0:
        s.flush_stats()
        assert abs(s.sampled_oct_per_ray_ - (512 * .9 + ref["n_oct"].mean() * .1)) < 1e-2


def test_compact_equals_dense():
    rig = load_rig("rig8")
    s = make_sampler(rig, mode=1)
    o, d, _ = rays_for(rig, 1500, seed=9)
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    world, warp, dirs, dists, ts, anchors, start_end, first = s.GetSamples(to, td)
    cs = s.sample_compact(to, td)
    V = int(cs.total.item())
    counts = cs.counts.cpu().numpy()
    assert V == counts.sum() == int(start_end[-1, 1].item())
    offs = cs.offsets.cpu().numpy()
    assert np.array_equal(offs[:-1], start_end[:, 0].cpu().numpy()) and offs[-1] == V
    m = torch.from_numpy(valid_mask(counts)).cuda()
    assert torch.equal(cs.pts01[:V], (warp[m] + 1.5) / 3.0)
    assert torch.equal(cs.t[:V], ts[m]) and torch.equal(cs.delta[:V], dists[m])
    assert torch.equal(cs.anchor[:V].long(), anchors[m][:, 0]) and torch.equal(cs.node[:V].long(), anchors[m][:, 1])
    ray = torch.arange(1500, device="cuda").unsqueeze(1).expand(-1, 1024)[m]
    assert torch.equal(cs.ray_id[:V].long(), ray)


def test_rays_that_miss_and_empty_batch():
    rig = load_rig("rig8")
    s = make_sampler(rig, mode=1)
    # rays far outside the root box pointing away: no intersection at all (the reference reads OOB here)
    o = np.array([[600., 600., 600.], [0.3, 0.2, 2.]], np.float32)
    d = np.array([[1., 0., 0.], [0., 0., -1.]], np.float32)
    out = s.GetSamples(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
    se = out[6].cpu().numpy()
    assert se[0, 1] - se[0, 0] == 0 and se[1, 1] - se[1, 0] > 0
    assert float(out[7][0]) == np.float32(1e9)
    ref = orc.sampler_get_samples(o, d, np.ones(1024 + 12, np.float32), rig["tree_nodes"], rig["pers_trans"])
    assert np.array_equal(se[:, 1] - se[:, 0], ref["counts"])
    cs = s.sample_compact(torch.zeros((0, 3), device="cuda"), torch.zeros((0, 3), device="cuda"))
    assert int(cs.total.item()) == 0


def test_update_oct_nodes_matches_oracle():
    rig = load_rig("rig8")
    s = make_sampler(rig, mode=1)
    R = 1200
    o, d, _ = rays_for(rig, R, seed=21)
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    out = s.GetSamples(to, td)
    anchors, start_end = out[5], out[6]
    rng = np.random.RandomState(0)
    w = (rng.rand(R, 1024, 1) ** 6).astype(np.float32) * 0.05
    a = (rng.rand(R, 1024, 1) ** 6).astype(np.float32) * 0.1
    counts = (start_end[:, 1] - start_end[:, 0]).cpu().numpy().astype(np.int32)
    nodes = rig["tree_nodes"].copy()
    n_nodes = nodes.size // 128
    ws, as_, vc = np.full(n_nodes, 1000, np.int64), np.full(n_nodes, 1000, np.int64), np.zeros(n_nodes, np.int64)
    ws[::7] = 0            # some nodes about to be pruned
    s.tree_weight_stats_.copy_(torch.from_numpy(ws))
    orc.update_oct_nodes(counts, anchors[..., 1].cpu().numpy().reshape(-1), w.reshape(-1), a.reshape(-1), nodes, ws, as_, vc)
    s.UpdateOctNodes(anchors, start_end.unsqueeze(1).expand(-1, 1024, -1), torch.from_numpy(w).cuda(),
                     torch.from_numpy(a).cuda(), 7)     # step 7: no milestone, no compaction
    assert np.array_equal(s.tree_weight_stats_.cpu().numpy(), ws)
    assert np.array_equal(s.tree_alpha_stats_.cpu().numpy(), as_)
    assert np.array_equal(s.tree_visit_cnt_.cpu().numpy(), vc)
    assert np.array_equal(s.tree_nodes_gpu_.cpu().numpy(), nodes)   # pruned leaves (trans_idx = -1) bit-exact
    assert (nodes.view(np.int64).reshape(-1, 16)[:, 12] != rig["tree_nodes"].view(np.int64).reshape(-1, 16)[:, 12]).sum() > 0


def test_trans_query_frame_matches_oracle():
    rig = load_rig("rig8")
    s = make_sampler(rig, mode=1)
    n_nodes = rig["tree_nodes"].size // 128
    rng = np.random.RandomState(1)
    anchors = rng.randint(-2, n_nodes + 2, size=4000).astype(np.int64)
    pts = rng.uniform(-3, 3, size=(4000, 3)).astype(np.float32)
    pts[:, 2] = rng.uniform(-1, 1, size=4000)
    ref = orc.trans_query_frame(rig["tree_nodes"], rig["pers_trans"], anchors, pts)
    got = s.TransQueryFrame(torch.from_numpy(pts).cuda(), torch.from_numpy(anchors).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
    print("trans_query_frame bit-equal fraction", np.mean(got == ref))


def test_proc_octree_roundtrip_and_resample():
    """compact + subdivide on the host keeps the device path consistent: after ProcOctree the CUDA sampler and the
    oracle still agree bit-exactly on the NEW blobs, and pruned leaves are gone."""
    rig = load_rig("rig8")
    s = make_sampler(rig, mode=1)
    R = 600
    o, d, _ = rays_for(rig, R, seed=33)
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    out = s.GetSamples(to, td)
    w = torch.zeros((R, 1024, 1), device="cuda")
    w[:, :200] = 0.5                                     # near samples "occupied", far ones empty
    for it in range(3):
        s.UpdateOctNodes(out[5], out[6].unsqueeze(1).expand(-1, 1024, -1), w, w, 7)
    n_before = s.n_nodes
    s.ProcOctree(True, True, False)
    assert s.n_nodes != n_before
    out2 = s.GetSamples(to, td)
    ref = orc.sampler_get_samples(o, d, np.ones(1024 + R + 10, np.float32), s.tree_nodes_gpu_.cpu().numpy(),
                                  s.pers_trans_gpu_.cpu().numpy())
    counts = (out2[6][:, 1] - out2[6][:, 0]).cpu().numpy()
    assert np.array_equal(counts, ref["counts"])
    m = valid_mask(counts)
    assert np.array_equal(out2[5].cpu().numpy()[m], ref["anchors"][m])


def test_cold_queries_match_oracle():
    """get_points_anchors (GetRaysTreeNodesIntersects + GetTreeNodeIdxFromTs, PersSampler_cuda.cu:799-853, 924-980)
    and GetEdgeSamples (:479-516) against the oracle, through the Python core."""
    import ctypes as C
    rig = load_rig("rig8")
    s = make_sampler(rig, mode=1)
    from gfnerf_b200.persoctree import rig_rays
    R, S = 200, 48
    o, d, _ = rig_rays(rig["c2w"], rig["intri"], R, seed=12)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    rng = np.random.RandomState(3)
    t0 = np.sort(rng.uniform(0.05, 12.0, size=(R, S + 1)).astype(np.float32), axis=1)
    ts, te = t0[:, :-1, None].copy(), t0[:, 1:, None].copy()
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    got = s.GetPointsAnchors(T(o), T(d.astype(np.float32)), T(ts), T(te))
    assert got.shape == (R, S, 1) and got.dtype == torch.int64
    t_cur = ((torch.from_numpy(ts) + torch.from_numpy(te)) / 2.0).numpy()[:, :, 0]
    ref = orc.points_anchors(o, d.astype(np.float32), t_cur, rig["tree_nodes"])
    assert np.array_equal(got.cpu().numpy()[:, :, 0], ref)
    assert (ref >= 0).mean() > 0.3                                     # most samples do fall into a leaf
    # every anchor is a leaf whose box contains the sample (up to fp slack)
    nodes = rig["tree_nodes"].view(np.uint8).reshape(-1, 128)
    cs = nodes[:, :16].copy().view(np.float32)
    m = ref >= 0
    p = o[:, None, :] + d[:, None, :] * t_cur[:, :, None]
    box = cs[ref[m]]
    assert np.all(np.abs(p[m] - box[:, :3]).max(-1) <= box[:, 3] * 0.5 * (1 + 1e-3) + 1e-4)
    # edge samples with given draws
    pool = s.edge_pool()
    n_edges = pool.numel() // 64
    assert n_edges > 0
    n = 500
    eidx = rng.randint(0, n_edges, size=n).astype(np.int64)
    ecoord = rng.uniform(-1, 1, size=(n, 2)).astype(np.float32)
    pts, idx = s.GetEdgeSamples(n, T(eidx), T(ecoord))
    rpts, ridx = orc.edge_samples(pool.cpu().numpy(), rig["pers_trans"], eidx, ecoord)
    assert np.array_equal(idx.cpu().numpy(), ridx) and np.array_equal(pts.cpu().numpy(), rpts)
    pts2, idx2 = s.GetEdgeSamples(64)                                  # random draws: shapes, valid transform ids
    assert pts2.shape == (64, 2, 3) and int(idx2.min()) >= 0 and int(idx2.max()) < rig["pers_trans"].size // 576
    assert torch.isfinite(pts2).all()
