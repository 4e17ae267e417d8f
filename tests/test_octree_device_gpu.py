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
