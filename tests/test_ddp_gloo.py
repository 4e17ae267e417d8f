"""The N > 1 host logic (gfnerf_b200.ddp) with world_size 2 over gloo on CPU: gradient SUM all-reduce through the
flat bucket, MAX all-reduce of octree votes, parameter broadcast, ray sharding.  The GPU box runs the same code over
NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gfnerf_b200.ddp import FlatBucket, GradSync, rank_seed, shard_slice
        sync = GradSync(dist.group.WORLD, torch.device("cpu"))
        assert sync.world == world and sync.comm_stream is None
        # parameters: rank 0's values everywhere
        g = torch.Generator().manual_seed(100 + rank)
        table = torch.rand(4096, 2, generator=g)
        mlp = torch.rand(11603, generator=g)
        sync.broadcast_([table, mlp])
        # gradients: MLP + embedding share one flat bucket, the table gradient is its own message
        bucket = FlatBucket([(11603,), (7, 32)])
        g_mlp, g_emb = bucket.views
        g_mlp.fill_(rank + 1.0)
        g_emb.copy_(torch.arange(7 * 32, dtype=torch.float32).view(7, 32) * (rank + 1))
        g_table = torch.full((4096, 2), float(10 ** rank))
        sync.start_sum([bucket.flat, g_table])
        sync.wait()
        # octree votes: adders start at -1, a rank that saw the leaf occupied votes 512 / 32, marks are 0 / 1
        n_nodes = 9
        scratch = torch.full((3 * n_nodes,), -1, dtype=torch.int64)
        scratch[2 * n_nodes:] = 0
        visit = torch.zeros(n_nodes, dtype=torch.int64)
        scratch[rank] = 512                      # leaf `rank` got a weight vote on this rank only
        scratch[n_nodes + 3] = 32 if rank == 1 else -1
        scratch[2 * n_nodes + rank] = 1
        visit[4] = 5 + 10 * rank
        sync.max_([scratch, visit])
        rays = np.arange(8192 * 3 + 1)
        sl = shard_slice(rays.size, rank, world)
        torch.save(dict(table=table, mlp=mlp, g_mlp=g_mlp.clone(), g_emb=g_emb.clone(), g_table=g_table,
                        scratch=scratch, visit=visit, seed=rank_seed(1234, rank), lo=sl.start, hi=sl.stop),
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"rank{k}.pt")) for k in range(world)]
    # broadcast: identical parameters
    assert torch.equal(r[0]["table"], r[1]["table"]) and torch.equal(r[0]["mlp"], r[1]["mlp"])
    # SUM over ranks, identical on every rank
    for k in range(world):
        assert torch.all(r[k]["g_mlp"] == 3.0)
        assert torch.equal(r[k]["g_emb"], torch.arange(7 * 32, dtype=torch.float32).view(7, 32) * 3)
        assert torch.all(r[k]["g_table"] == 11.0)
    # MAX over ranks of the votes: both leaves voted, alpha vote of leaf 3, both marks, larger visit count
    for k in range(world):
        s, n = r[k]["scratch"], 9
        assert s[0] == 512 and s[1] == 512 and s[2] == -1
        assert s[n + 3] == 32 and s[n + 2] == -1
        assert s[2 * n] == 1 and s[2 * n + 1] == 1 and s[2 * n + 2] == 0
        assert r[k]["visit"][4] == 15
    # rays: distinct seeds, a balanced disjoint cover
    assert r[0]["seed"] != r[1]["seed"]
    assert r[0]["lo"] == 0 and r[0]["hi"] == r[1]["lo"] and r[1]["hi"] == 8192 * 3 + 1
    assert abs((r[0]["hi"] - r[0]["lo"]) - (r[1]["hi"] - r[1]["lo"])) <= 1


def test_single_process_is_a_no_op():
    from gfnerf_b200.ddp import GradSync, shard_slice
    sync = GradSync(None, torch.device("cpu"))
    assert sync.world == 1
    t = torch.ones(4)
    sync.start_sum([t])
    sync.wait()
    sync.max_([t])
    assert torch.all(t == 1)
    assert shard_slice(10, 0, 1) == slice(0, 10)
    covered = sum((list(range(*shard_slice(10, k, 4).indices(10))) for k in range(4)), [])
    assert covered == list(range(10))


def test_table_reduce_ranges_cover_the_reachable_rows_once():
    from gfnerf_b200.ddp import table_level_rows, table_reduce_ranges
    T = 1 << 10
    assert table_level_rows(T, 0) == (0, T) and table_level_rows(T, 1) == (T // 2, T // 2 + T)   # half-overlapping
    for group in (1, 2, 4, 8, 16):
        rr = table_reduce_ranges(T, group)
        assert rr[0][2] == 0 and rr[-1][3] == 15 * T // 2 + T == 17 * T // 2
        assert all(a[3] == b[2] and a[1] == b[0] for a, b in zip(rr, rr[1:]))                    # contiguous, disjoint
        for l0, l1, lo, hi in rr:
            # nothing a LATER level can still write lies inside a range that is handed over
            assert all(table_level_rows(T, l)[0] >= hi for l in range(l1, 16))
            # and everything levels < l1 wrote that was not handed over earlier is inside or comes later
            assert lo == table_level_rows(T, l0)[0]
    assert table_reduce_ranges(48 * 16, 16) == [(0, 16, 0, 17 * 48 * 8)]


def _worker_groups(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gfnerf_b200.ddp import GradSync, table_level_rows, table_reduce_ranges
        sync = GradSync(dist.group.WORLD, torch.device("cpu"))
        T = 256
        rng = np.random.RandomState(7 + rank)
        contrib = []                                   # per level: (rows, values) this rank scatters
        for l in range(16):
            lo, hi = table_level_rows(T, l)
            contrib.append((torch.from_numpy(rng.randint(lo, hi, size=300)), torch.from_numpy(rng.normal(size=(300, 2)))))
        results = {}
        for group in (16, 4, 1):
            g = torch.zeros(16 * T, 2, dtype=torch.float64)
            for l0, l1, r0, r1 in table_reduce_ranges(T, group):
                for l in range(l0, l1):                # the scatter of this level group
                    g.index_add_(0, contrib[l][0], contrib[l][1])
                sync.start_sum([g[r0:r1]])             # exactly what GFNeRFEngine.train_step hands over
            sync.wait()
            results[group] = g
        torch.save(results, os.path.join(out_dir, f"groups{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_level_group_exchange_equals_one_shot_exchange(tmp_path):
    """With the reference's half-overlapping level windows a level group is no longer a private row range; the ranges
    of table_reduce_ranges still give every rank the same total as one all-reduce after the whole scatter."""
    world = 2
    mp.spawn(_worker_groups, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"groups{k}.pt")) for k in range(world)]
    assert torch.equal(r[0][16], r[1][16]) and float(r[0][16].abs().sum()) > 0
    assert not r[0][16][17 * 256 // 2:].any()                       # rows past 8.5 T: never touched
    for group in (4, 1):
        for k in range(world):
            assert torch.allclose(r[k][group], r[0][16], rtol=0, atol=1e-12)
