"""The TorchScript custom classes `torch.classes.my_classes.{Hash3DAnchored, PersSampler}` (csrc/torch_bindings.cpp,
the reference's own operator boundary, gfnerf/bindings/hashanchored/bindings.cpp:343-401) loaded the way the
reference loads its f2nerf-bindings.so, against the Python host layer over the same C-ABI."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gf-nerf_b200", "f2nerf_bindings_b200.so")


def load():
    if not os.path.exists(SO):
        import importlib.util
        spec = importlib.util.spec_from_file_location("gf_build", os.path.join(ROOT, "gf-nerf_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
        mod.build_torch_bindings()
    torch.classes.load_library(SO)      # gfnerf/hash_3d_anchored.py:13-14
    return torch.classes.my_classes


def test_library_registers_the_reference_class_names():
    mc = load()
    s = mc.PersSampler()
    for name in ("InitSampler", "GetSamples", "UpdateOctNodes", "UpdateRayMarch", "UpdateMode", "States", "LoadStates",
                 "trans_query_frame", "get_ray_march_fineness_", "get_n_volumes_"):
        assert hasattr(s, name), name
    s.UpdateMode(1)
    assert s.get_mode_() == 1
    s.Configure(1000, 1024, 0.01, True, 1.0 / 256, 0, 1.0, 16.0, 10000)
    s.UpdateRayMarch(0)
    assert abs(s.get_ray_march_fineness_() - 16.0) < 1e-5
    s.UpdateRayMarch(10000)
    assert s.get_ray_march_fineness_() == 1.0
    for name in ("ProcOctree", "MarkInvisibleNodes", "qurey_tree_nodes_centers", "UpdateBlockIdxs", "VisOctree",
                 "get_sub_div_milestones_", "get_tree_nodes_center_", "get_tree_nodes_side_len_",
                 "get_tree_nodes_is_leaf_node_", "get_tree_nodes_trans_idx_", "get_tree_nodes_block_idx_",
                 "get_pers_trans_info", "get_points_anchors", "GetEdgeSamples"):
        assert hasattr(s, name), name
    assert hasattr(mc, "Hash3DAnchored")


@pytest.mark.gpu
def test_hash3d_class_equals_python_core():
    from gfnerf_b200.hash_3d_anchored import Hash3DAnchoredCore
    from tests.helpers import hash_inputs
    mc = load()
    n, n_vol, log2T = 5000, 6, 13
    _, _, _, pts, anchors = hash_inputs(n, n_vol, log2T, seed=4)
    py = Hash3DAnchoredCore(log2T, n_vol)
    py.Reset()
    cc = mc.Hash3DAnchored(log2T, n_vol, 0.1)
    assert cc.LoadStates(py.States(), 0) == 4
    for a, b in zip(cc.States(), py.States()):
        assert torch.equal(a.cpu(), b.cpu())
    tp, ta = torch.from_numpy(pts).cuda(), torch.from_numpy(anchors).cuda()
    out_c = cc.AnchoredQuery(tp, ta)
    out_p = py.AnchoredQuery(tp, ta)
    assert out_c.shape == (n, 32) and torch.equal(out_c, out_p)                      # bit-exact encodings
    g = torch.randn(n, 32, generator=torch.Generator().manual_seed(0)).cuda() * 1e-3
    out_c.backward(g)
    out_p.backward(g)
    gc, gp = cc.GetParams()[0].grad, py.GetParams()[0].grad
    assert float((gc - gp).abs().max()) <= 1e-5 * float(gp.abs().max())            # fp32 atomics: order only
    cc.Zero()
    assert float(cc.AnchoredQuery(tp, ta).abs().max()) == 0.0
    cc.SetFeatPoolRequireGrad(False)
    assert not cc.AnchoredQuery(tp, ta).requires_grad


@pytest.mark.gpu
def test_perssampler_class_equals_python_core():
    from gfnerf_b200.persoctree import rig_rays
    from tests.helpers import load_rig, make_sampler
    mc = load()
    rig = load_rig("rig8")
    py = make_sampler(rig, mode=1)
    cc = mc.PersSampler()
    cc.Configure(1000, 1024, 0.01, True, 1.0 / 256, 1, 1.0, 16.0, 10000)
    assert cc.LoadStates(py.States(), 0) == 4
    assert cc.get_n_volumes_() == py.get_n_volumes()
    o, d, _ = rig_rays(rig["c2w"], rig["intri"], 300, seed=8)
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    bounds = torch.zeros(300, 2).cuda()
    a, b = cc.GetSamples(to, td, bounds), py.GetSamples(to, td, bounds)
    assert len(a) == 8
    for x, y in zip(a, b):
        assert x.shape == y.shape and x.dtype == y.dtype and torch.equal(x, y)     # bit-exact dense tensors
    w = torch.rand(300, 1024, 1).cuda() * (a[3] > 0).unsqueeze(-1)
    se = a[6].unsqueeze(1).expand(-1, 1024, -1).contiguous()
    cc.UpdateOctNodes(a[5], se, w, w, 10)
    py._vote(300, (a[6][:, 1] - a[6][:, 0]).int().contiguous(),
             (torch.arange(301, device="cuda") * 1024).int(), a[5].reshape(-1, 3)[:, 1].int().contiguous(),
             w.reshape(-1).contiguous(), w.reshape(-1).contiguous())
    assert torch.equal(cc.States()[2], py.States()[2])                               # visit counts
    wp = torch.rand(50, 3).cuda()
    nodes = a[5][0, :50, 1].contiguous()
    assert torch.equal(cc.trans_query_frame(wp, nodes), py.TransQueryFrame(wp, nodes))


@pytest.mark.gpu
def test_init_sampler_builds_the_fixture_octree_and_maintains_it():
    """InitSampler with the reference's argument list (gfnerf/perssampler.py:103-123): the octree comes from the C++
    host builder; then a vote with a subdivision milestone due (UpdateOctNodes -> ProcOctree / MarkInvisibleNodes)
    against the Python host layer fed with the same blobs."""
    from gfnerf_b200.persoctree import rig_rays
    from tests.helpers import load_rig, make_sampler
    mc = load()
    rig = load_rig("rig8")
    c2w = torch.from_numpy(rig["c2w"])
    w2c4 = torch.eye(4).repeat(c2w.shape[0], 1, 1)
    w2c4[:, :3, :] = c2w
    w2c = torch.linalg.inv(w2c4)[:, :3, :].contiguous()
    cc = mc.PersSampler()
    cc.InitSampler(1.5, [3, 4000], 1000, 1024, 0.01, True, 10, 1.0 / 256, 16, c2w, w2c, torch.from_numpy(rig["intri"]),
                   torch.from_numpy(rig["bounds"]), 1, 512, 1.0, 16.0, 10000)
    st = cc.States()
    assert torch.equal(st[0].cpu(), torch.from_numpy(rig["tree_nodes"]))          # node blob: the fixture, byte for byte
    assert st[1].numel() == rig["pers_trans"].size and st[3].cpu().tolist() == [4000, 3]
    # same blobs into the Python layer; one vote at step 3 = first milestone: subdivide, prune invisible, compact
    py = make_sampler(rig, mode=1)
    py.LoadStates([s.clone() for s in st], 0)
    assert py.sub_div_milestones_ == [4000, 3]
    o, d, _ = rig_rays(rig["c2w"], rig["intri"], 400, seed=3)
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    bounds = torch.zeros(400, 2).cuda()
    a, b = cc.GetSamples(to, td, bounds), py.GetSamples(to, td, bounds)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    w = torch.rand(400, 1024, 1, generator=torch.Generator().manual_seed(1)).cuda() * (a[3] > 0).unsqueeze(-1)
    se = a[6].unsqueeze(1).expand(-1, 1024, -1).contiguous()
    cc.UpdateOctNodes(a[5], se, w, w, 3)
    py.UpdateOctNodes(a[5], se, w, w, 3)
    sc, sp = cc.States(), py.States()
    assert sc[0].numel() > st[0].numel()                                            # leaves were split
    assert torch.equal(sc[0], sp[0]) and torch.equal(sc[2], sp[2])
    assert sc[3].cpu().tolist() == [4000] == sp[3].cpu().tolist()
    a2, b2 = cc.GetSamples(to, td, bounds), py.GetSamples(to, td, bounds)
    for x, y in zip(a2, b2):
        assert torch.equal(x, y)


@pytest.mark.gpu
def test_cold_surface_equals_python_core(tmp_path):
    """The introspection / cold methods the reference's Python wrapper calls (gfnerf/perssampler.py:448-640):
    getters as nested lists, get_pers_trans_info, qurey_tree_nodes_centers (sic), UpdateBlockIdxs, VisOctree."""
    from tests.helpers import load_rig, make_sampler
    mc = load()
    rig = load_rig("rig8")
    py = make_sampler(rig, mode=1)
    cc = mc.PersSampler()
    cc.Configure(1000, 1024, 0.01, True, 1.0 / 256, 1, 1.0, 16.0, 10000)
    cc.LoadStates(py.States(), 0)
    assert cc.get_sub_div_milestones_() == py.sub_div_milestones_
    assert np.allclose(np.array(cc.get_tree_nodes_center_(), np.float32), np.array(py.get_tree_nodes_center_(), np.float32))
    assert cc.get_tree_nodes_side_len_() == py.get_tree_nodes_side_len_()
    assert cc.get_tree_nodes_trans_idx_() == py.get_tree_nodes_trans_idx_()
    assert cc.get_tree_nodes_block_idx_() == py.get_tree_nodes_block_idx_()
    assert list(cc.get_tree_nodes_is_leaf_node_()) == py.get_tree_nodes_is_leaf_node_()
    w2xz, weight, center, side, dis = cc.get_pers_trans_info()
    tr = py.pers_trans_gpu_.cpu().numpy().view(np.float32).reshape(-1, 144)
    assert np.array_equal(np.array(w2xz, np.float32).reshape(-1, 96), tr[:, :96])
    assert np.array_equal(np.array(weight, np.float32).reshape(-1, 36), tr[:, 96:132])
    assert np.array_equal(np.array(center, np.float32), tr[:, 132:135])
    assert np.array_equal(np.array(side, np.float32), tr[:, 135]) and np.array_equal(np.array(dis, np.float32), tr[:, 136])
    anchors = torch.tensor([[0], [5], [17], [10 ** 9], [-1]], dtype=torch.int64).cuda()
    got = cc.qurey_tree_nodes_centers(anchors)
    ref = py.QueryTreeNodeCenters(anchors[:3, 0])
    assert torch.equal(got[:3], ref) and not got[3:].any()                # out-of-range anchors leave zeros (:998)
    from gfnerf_b200.persoctree import rig_rays
    o, d, _ = rig_rays(rig["c2w"], rig["intri"], 64, seed=2)
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    t0 = torch.sort(torch.rand(64, 33, generator=torch.Generator().manual_seed(0)) * 10 + 0.05, dim=1).values
    ts, te = t0[:, :-1, None].contiguous().cuda(), t0[:, 1:, None].contiguous().cuda()
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    assert torch.equal(cc.get_points_anchors(to, td, ts, te), py.GetPointsAnchors(to, td, ts, te))
    epts, eidx = cc.GetEdgeSamples(32)
    assert epts.shape == (32, 2, 3) and eidx.shape == (32, 2) and torch.isfinite(epts).all()
    assert int(eidx.min()) >= 0 and int(eidx.max()) < py.pers_trans_gpu_.numel() // 576
    centers = torch.tensor([[-2.0, -2.0, 0.0], [2.0, 2.0, 0.0], [2.0, -2.0, 0.0]]).cuda()
    cc.UpdateBlockIdxs(centers)
    py.UpdateBlockIdxs(centers)
    assert torch.equal(cc.States()[0], py.States()[0])
    assert set(cc.get_tree_nodes_block_idx_()) <= {0, 1, 2}
    cc.VisOctree(str(tmp_path))
    lines = (tmp_path / "octree.obj").read_text().splitlines()
    n = len(cc.get_tree_nodes_side_len_())
    assert sum(l.startswith("v ") for l in lines) == 8 * n
    assert sum(l.startswith("l ") for l in lines) == 12 * sum(cc.get_tree_nodes_is_leaf_node_())
