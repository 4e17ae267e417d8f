"""GPU parity: gf_octree_proc_device (csrc/octree_device.cu -- PersOctree::ProcOctree, PtsSampler/PersSampler.cpp:154-417,
rebuilt in HBM by one kernel) against gf_octree_proc, the host C++ restatement that tests/test_octree_host.py and
tests/test_ref_kernels.py pin byte for byte to the numpy restatement and to the reference's own ProcOctree body.
Bar: identical node blobs (all 128 bytes of every node) and identical statistics, for every flag combination, on trees
with pruning histories, through repeated subdivision, and the same error behaviour."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.helpers import load_rig, make_sampler
from tests.test_octree_host import _cxx, _octree

pytestmark = pytest.mark.gpu


def _dev(oc, compact, subdivide, brute):
    """-> (nodes uint8, weight_stats, alpha_stats, error word) from the device kernel"""
    from gfnerf_b200 import _lib
    L = _lib.lib()
    nodes = torch.from_numpy(np.ascontiguousarray(oc.nodes).view(np.uint8).reshape(-1).copy()).cuda()
    w, a, v = (torch.from_numpy(np.ascontiguousarray(x, np.int64)).cuda() for x in
               (oc.weight_stats, oc.alpha_stats, oc.visit_cnt))
    n_in = nodes.numel() // 128
    cap = 9 * n_in if subdivide else n_in
    sb = int(L.gf_octree_proc_device_scratch_bytes(n_in))
    scratch = torch.empty(sb, dtype=torch.uint8, device="cuda")
    o_nodes = torch.full((cap * 128,), 0xAB, dtype=torch.uint8, device="cuda")      # every output byte must be written
    o_w, o_a = torch.empty(cap, dtype=torch.int64, device="cuda"), torch.empty(cap, dtype=torch.int64, device="cuda")
    res = torch.zeros(4, dtype=torch.int32, device="cuda")
    _lib.check(L.gf_octree_proc_device(_lib.ptr(nodes), n_in, _lib.ptr(w), _lib.ptr(a), _lib.ptr(v), int(compact),
                                       int(subdivide), int(brute), _lib.ptr(o_nodes), _lib.ptr(o_w), _lib.ptr(o_a), cap,
                                       _lib.ptr(scratch), sb, _lib.ptr(res), res.data_ptr() + 8, _lib.cur_stream()))
    h = res.cpu()
    n, err = int(h[:2].view(torch.int64).item()), int(h[2].item())
    return o_nodes[:n * 128].cpu().numpy(), o_w[:n].cpu().numpy(), o_a[:n].cpu().numpy(), err


def _history(oc, rng, frac):
    n = oc.nodes.shape[0]
    leaves = np.nonzero(oc.nodes["trans_idx"] >= 0)[0]
    dead = rng.choice(leaves, size=int(len(leaves) * frac), replace=False)
    oc.nodes["trans_idx"][dead] = -1
    oc.weight_stats = rng.randint(-100, 5000, size=n).astype(np.int64)
    oc.alpha_stats = rng.randint(-100, 5000, size=n).astype(np.int64)
    oc.visit_cnt = rng.randint(0, 12, size=n).astype(np.int64)


@pytest.mark.parametrize("rig_name", ["rig8", "rig20"])
@pytest.mark.parametrize("compact,subdivide,brute", [(True, False, False), (True, True, False), (True, True, True),
                                                      (False, True, False), (False, False, False)])
def test_device_proc_equals_host_proc(rig_name, compact, subdivide, brute):
    rig = load_rig(rig_name)
    rng = np.random.RandomState(5)
    for trial in range(3):
        oc = _octree(rig)
        _history(oc, rng, 0.2 + 0.3 * trial)          # up to 80 % of the leaves voted empty: long single-child chains
        if not compact:
            # pruned leaves stay linked: the reference CHECK-fails, the host restatement raises, the kernel reports 2
            *_, err = _dev(oc, compact, subdivide, brute)
            assert err & 2
            oc.proc_octree(True, False, False)
            oc.visit_cnt = rng.randint(0, 12, size=oc.nodes.shape[0]).astype(np.int64)
        ref_nodes, ref_w, ref_a = _cxx(oc, compact, subdivide, brute)
        got_nodes, got_w, got_a, err = _dev(oc, compact, subdivide, brute)
        assert err == 0
        assert got_nodes.size == ref_nodes.size, (got_nodes.size // 128, ref_nodes.size // 128)
        assert np.array_equal(got_nodes, ref_nodes)
        assert np.array_equal(got_w, ref_w) and np.array_equal(got_a, ref_a)


def test_repeated_subdivision_and_pruning_stays_identical():
    """the milestone sequence (PersSampler.cpp:657-677): subdivide + compact, prune, compact ... three rounds deep"""
    rig = load_rig("rig8")
    rng = np.random.RandomState(11)
    oc = _octree(rig)
    for rnd in range(3):
        oc.visit_cnt = rng.randint(0, 12, size=oc.nodes.shape[0]).astype(np.int64)
        for flags in ((True, True, rnd == 0), (True, False, False)):
            ref_nodes, ref_w, ref_a = _cxx(oc, *flags)
            got_nodes, got_w, got_a, err = _dev(oc, *flags)
            assert err == 0 and np.array_equal(got_nodes, ref_nodes)
            assert np.array_equal(got_w, ref_w) and np.array_equal(got_a, ref_a)
            oc.nodes = ref_nodes.view(oc.nodes.dtype).copy()
            oc.weight_stats, oc.alpha_stats = ref_w, ref_a
            oc.visit_cnt = np.zeros(oc.nodes.shape[0], np.int64)
            if flags[1]:                                  # votes between the two passes of a milestone
                leaves = np.nonzero(oc.nodes["trans_idx"] >= 0)[0]
                oc.nodes["trans_idx"][rng.choice(leaves, size=len(leaves) // 3, replace=False)] = -1
    assert oc.nodes.shape[0] > rig["tree_nodes"].size // 128


def test_everything_pruned_leaves_the_bare_root_like_the_host():
    """every leaf voted empty: the pruning sweep starts at node 1 (PersSampler.cpp:189), so the root stays an interior
    node without children and is the one survivor -- on the host and on the device"""
    rig = load_rig("rig8")
    oc = _octree(rig)
    oc.nodes["trans_idx"][:] = -1
    ref_nodes, ref_w, ref_a = _cxx(oc, True, False, False)
    got_nodes, got_w, got_a, err = _dev(oc, True, False, False)
    assert err == 0 and ref_nodes.size == 128
    assert np.array_equal(got_nodes, ref_nodes) and np.array_equal(got_w, ref_w) and np.array_equal(got_a, ref_a)


def test_sampler_proc_octree_runs_on_the_device_and_matches_the_host_schedule():
    """PersSamplerCore.ProcOctree (device) against ProcOctreeHost (the reference's D2H / host / H2D schedule) on a
    sampler that has been trained on: same blobs, same statistics, and sampling goes on identically afterwards."""
    from gfnerf_b200.persoctree import rig_rays
    rig = load_rig("rig8")
    a, b = make_sampler(rig, mode=1), make_sampler(rig, mode=1)
    rng = np.random.RandomState(3)
    n = a.n_nodes
    nodes = a.tree_nodes_gpu_.view(-1, 128).clone()
    tidx = nodes[:, 96:104].contiguous().view(torch.int64).view(-1)
    leaves = torch.nonzero(tidx >= 0).view(-1)
    tidx[leaves[torch.from_numpy(rng.choice(len(leaves), size=len(leaves) // 3, replace=False)).cuda()]] = -1
    nodes[:, 96:104] = tidx.view(-1, 1).view(torch.uint8)
    visit = torch.from_numpy(rng.randint(0, 12, size=n).astype(np.int64)).cuda()
    for s in (a, b):
        s.tree_nodes_gpu_ = nodes.reshape(-1).clone()
        s.tree_visit_cnt_ = visit.clone()
    for flags in ((True, True, False), (True, False, False)):
        a.ProcOctree(*flags)
        b.ProcOctreeHost(*flags)
        assert torch.equal(a.tree_nodes_gpu_, b.tree_nodes_gpu_)
        assert torch.equal(a.tree_weight_stats_, b.tree_weight_stats_) and torch.equal(a.tree_alpha_stats_, b.tree_alpha_stats_)
        assert torch.equal(a.tree_visit_cnt_, b.tree_visit_cnt_) and a.n_nodes == b.n_nodes
    assert np.array_equal(a.octree.tree_nodes_blob(), b.octree.tree_nodes_blob())      # the lazily refreshed host mirror
    o, d, _ = rig_rays(rig["c2w"], rig["intri"], 256, seed=4)
    ca = a.sample_compact(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
    cb = b.sample_compact(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
    V = int(ca.total.item())
    assert V == int(cb.total.item()) and V > 0
    assert torch.equal(ca.node[:V], cb.node[:V]) and torch.equal(ca.t[:V], cb.t[:V])


# ---- MarkInvisibleNodes / UpdateBlockIdxs on the device (gf_octree_mark_invisible, gf_octree_set_block_idxs) ----------
def _cameras(rig, n_cams):
    c2w = rig["c2w"][:n_cams]
    m = np.tile(np.eye(4, dtype=np.float32)[None], (c2w.shape[0], 1, 1))
    m[:, :3, :] = c2w
    w2c = np.ascontiguousarray(np.linalg.inv(m)[:, :3, :].astype(np.float32)) if n_cams else np.zeros((0, 3, 4), np.float32)
    return w2c, np.ascontiguousarray(rig["intri"][:n_cams]), np.ascontiguousarray(rig["bounds"][:n_cams])


def _many_nodes(rig, copies, seed):
    """The rig's node blob `copies` times over, centres and side lengths jittered: tens of thousands of spheres at
    every distance from the frustum faces, so that the borderline comparisons of CheckVisible are exercised."""
    rng = np.random.RandomState(seed)
    blobs = [rig["tree_nodes"].reshape(-1, 128).copy() for _ in range(copies)]
    for k, b in enumerate(blobs[1:]):
        cs = b[:, :16].copy().view(np.float32)
        cs[:, :3] += rng.normal(0, 0.5 * (k + 1), size=cs[:, :3].shape).astype(np.float32)
        cs[:, 3] *= rng.uniform(0.25, 2.0, size=cs.shape[0]).astype(np.float32)
        b[:, :16] = cs.view(np.uint8)
    return np.ascontiguousarray(np.concatenate(blobs, 0).reshape(-1))


@pytest.mark.parametrize("rig_name,n_cams,copies", [("rig8", 6, 1), ("rig8", 64, 16), ("rig20", 3, 4), ("rig20", 400, 8),
                                                     ("rig8", 0, 1), ("rig8", 129, 3)])
def test_mark_invisible_kernel_equals_oracle_and_the_reference_kernel(rig_name, n_cams, copies):
    """Bar: the node blob the kernel leaves is byte-identical to the oracle's (PersSampler_cuda.cu:680-742 restated
    with nvcc's contraction) and to what the reference's own kernel, built by nvcc, leaves on the same B200."""
    from gfnerf_b200 import _lib
    from oracle import oracle as orc
    from oracle import ref_host as rh
    rig = load_rig(rig_name)
    if n_cams > rig["c2w"].shape[0]:                                    # more cameras than one shared-memory tile
        rig = dict(rig)
        reps = -(-n_cams // rig["c2w"].shape[0])
        for k in ("c2w", "intri", "bounds"):
            rig[k] = np.concatenate([rig[k]] * reps, 0)
        rig["c2w"] = rig["c2w"].copy()
        rig["c2w"][:, :3, 3] += np.random.RandomState(1).normal(0, 0.3, size=(rig["c2w"].shape[0], 3)).astype(np.float32)
    w2c, intri, bound = _cameras(rig, n_cams)
    blob = _many_nodes(rig, copies, seed=n_cams)
    want = orc.mark_invisible_nodes(blob, intri, w2c, bound)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    nodes, tw, ti, tb = T(blob), T(w2c), T(intri), T(bound)
    _lib.check(_lib.lib().gf_octree_mark_invisible(_lib.ptr(nodes), nodes.numel() // 128, _lib.ptr(tw), _lib.ptr(ti),
                                                   _lib.ptr(tb), n_cams, _lib.cur_stream()), "gf_octree_mark_invisible")
    got = nodes.cpu().numpy()
    tidx0 = blob.view(np.int64).reshape(-1, 16)[:, 12]
    marked = (got.view(np.int64).reshape(-1, 16)[:, 12] == -1) & (tidx0 != -1)
    if n_cams:
        assert 0 < marked.sum() < (tidx0 != -1).sum()
    else:
        assert (got.view(np.int64).reshape(-1, 16)[:, 12] == -1).all()
    assert np.array_equal(got, want), (got != want).sum()
    if rh.cuda_available() and n_cams:
        ref = T(blob)
        rh.cuda_mark_invisible_nodes(ref, ti, tw, tb)
        assert torch.equal(ref, nodes), int((ref != nodes).sum())


@pytest.mark.parametrize("n_blocks", [1, 7, 600])
def test_set_block_idxs_kernel_equals_oracle_and_the_reference_kernel(n_blocks):
    from gfnerf_b200 import _lib
    from oracle import oracle as orc
    from oracle import ref_host as rh
    rig = load_rig("rig20")
    blob = _many_nodes(rig, 4, seed=n_blocks)
    rng = np.random.RandomState(n_blocks)
    centers = rng.uniform(-4, 4, size=(n_blocks, 3)).astype(np.float32)
    if n_blocks >= 7:
        centers[5:7] = centers[1:3]                                    # exact ties: the first of equal minima wins
    want = orc.set_block_idxs(blob, centers)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    nodes, tc = T(blob), T(centers)
    _lib.check(_lib.lib().gf_octree_set_block_idxs(_lib.ptr(nodes), nodes.numel() // 128, _lib.ptr(tc), n_blocks,
                                                   _lib.cur_stream()), "gf_octree_set_block_idxs")
    got = nodes.cpu().numpy()
    idx = got.view(np.int64).reshape(-1, 16)[:, 13]
    assert idx.min() >= 0 and idx.max() < n_blocks and (n_blocks < 7 or not np.isin(idx, (5, 6)).any())
    assert np.array_equal(got, want), (got != want).sum()
    if rh.cuda_available():
        ref = T(blob)
        rh.cuda_set_block_idxs(ref, tc)
        assert torch.equal(ref, nodes), int((ref != nodes).sum())
    # nothing closer than 1e9: -1, like the reference's initial value
    far = T(np.full((2, 3), 3e9, np.float32))
    _lib.check(_lib.lib().gf_octree_set_block_idxs(_lib.ptr(nodes), nodes.numel() // 128, _lib.ptr(far), 2,
                                                   _lib.cur_stream()), "gf_octree_set_block_idxs")
    assert (nodes.cpu().numpy().view(np.int64).reshape(-1, 16)[:, 13] == -1).all()


def test_sampler_mark_invisible_nodes_and_update_block_idxs():
    """PersSamplerCore.MarkInvisibleNodes / UpdateBlockIdxs (the methods the milestones and the focal stage call) leave
    the blobs the oracle leaves; UpdateBlockIdxs compacts afterwards like the reference (:767-798)."""
    from oracle import oracle as orc
    rig = load_rig("rig8")
    s = make_sampler(rig, mode=1)
    keep = 5                                                            # a sampler that only knows five cameras
    s.w2c_, s.intri_, s.bound_ = s.w2c_[:keep].contiguous(), s.intri_[:keep].contiguous(), s.bound_[:keep].contiguous()
    before = s.tree_nodes_gpu_.cpu().numpy().copy()
    s.MarkInvisibleNodes()
    want = orc.mark_invisible_nodes(before, s.intri_.cpu().numpy(), s.w2c_.cpu().numpy(), s.bound_.cpu().numpy())
    assert np.array_equal(s.tree_nodes_gpu_.cpu().numpy(), want) and not np.array_equal(want, before)
    centers = np.array([[-2.0, -2.0, 0.0], [2.0, 2.0, 0.0], [2.0, -2.0, 0.0]], np.float32)
    marked = orc.set_block_idxs(want, centers)
    b = make_sampler(rig, mode=1)                                       # the same through the host schedule
    b.tree_nodes_gpu_ = torch.from_numpy(marked).cuda()
    b.ProcOctreeHost(True, False, False)
    s.UpdateBlockIdxs(torch.from_numpy(centers))
    assert torch.equal(s.tree_nodes_gpu_, b.tree_nodes_gpu_)
    assert set(s.get_tree_nodes_block_idx_()) <= {0, 1, 2}


@pytest.mark.parametrize("name", ["rig8", "rig20"])
def test_octree_mark_kernels_match_the_reference_fixture(name):
    """gf_octree_mark_invisible + gf_octree_set_block_idxs against tests/golden/ref_marks.npz: the outputs of the
    reference's own kernels (compiled for the host in the build container) on the same node blob, cameras, centres."""
    import os
    from gfnerf_b200 import _lib
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_marks.npz"))
    tree = load_rig(name)["tree_nodes"]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    nodes, tw, ti, tb, tc = T(tree), T(fx[f"{name}_w2c"]), T(fx[f"{name}_intri"]), T(fx[f"{name}_bounds"]), T(fx[f"{name}_centers"])
    L, n = _lib.lib(), tree.size // 128
    _lib.check(L.gf_octree_mark_invisible(_lib.ptr(nodes), n, _lib.ptr(tw), _lib.ptr(ti), _lib.ptr(tb), tw.shape[0],
                                          _lib.cur_stream()), "gf_octree_mark_invisible")
    _lib.check(L.gf_octree_set_block_idxs(_lib.ptr(nodes), n, _lib.ptr(tc), tc.shape[0], _lib.cur_stream()),
               "gf_octree_set_block_idxs")
    blob, before = nodes.cpu().numpy().view(np.int64).reshape(-1, 16), tree.view(np.int64).reshape(-1, 16)
    assert np.array_equal(blob[:, 12], fx[f"{name}_trans_idx"]) and np.array_equal(blob[:, 13], fx[f"{name}_block_idx"])
    keep = [c for c in range(16) if c not in (12, 13)]
    assert np.array_equal(blob[:, keep], before[:, keep])
