"""Shared synthetic-input builders for the parity tests (seeded, small)."""
import numpy as np


def first_primes_from(start, count):
    """`count` consecutive primes >= start (deterministic prime pool for fixtures)."""
    out = []
    x = start | 1
    while len(out) < count:
        i, ok = 3, x % 2 == 1
        while ok and i * i <= x:
            ok = x % i != 0
            i += 2
        if ok:
            out.append(x)
        x += 2
    return np.array(out, np.int64)


def lcg_primes(count, seed=1234):
    """primes in [2^28, 2^30) picked by a seeded LCG walk (SURVEY 8d fixture recipe)."""
    rng = np.random.RandomState(seed)
    out = []
    while len(out) < count:
        x = int(rng.randint(1 << 28, 1 << 30)) | 1
        while True:
            i, ok = 3, True
            while ok and i * i <= x:
                ok = x % i != 0
                i += 2
            if ok:
                break
            x += 2
        if x < (1 << 30):
            out.append(x)
    return np.array(out, np.int32)


def _is_prime_mr(x):
    """deterministic Miller-Rabin (bases 2, 7, 61: exact below 4.7e9), vectorised; x odd int64 array, 3 < x < 2^31"""
    x = x.astype(np.uint64)
    d = x - np.uint64(1)
    s = np.zeros(x.shape, np.int64)
    while True:
        even = (d & np.uint64(1)) == 0
        if not even.any():
            break
        d = np.where(even, d >> np.uint64(1), d)
        s += even
    res = np.ones(x.shape, bool)
    for a in (2, 7, 61):
        y = np.ones_like(x)
        base = np.full_like(x, a) % x
        e = d.copy()
        while e.any():
            odd = (e & np.uint64(1)) == 1
            y = np.where(odd, (y * base) % x, y)
            base = (base * base) % x
            e >>= np.uint64(1)
        ok = (y == 1) | (y == x - np.uint64(1))
        for r in range(1, int(s.max())):
            y = (y * y) % x
            ok |= (y == x - np.uint64(1)) & (r < s)
        res &= ok
    return res


def fast_primes(count, seed=1234):
    """`count` random primes in [2^28, 2^30) (fixture prime pools), seeded."""
    rng = np.random.RandomState(seed)
    out = np.empty(0, np.int64)
    while out.size < count:
        cand = rng.randint(1 << 28, 1 << 30, size=(count - out.size) * 12 + 64).astype(np.int64) | 1
        cand = cand[(cand % 3 != 0) & (cand % 5 != 0) & (cand % 7 != 0)]
        out = np.concatenate([out, cand[_is_prime_mr(cand)]])
    return out[:count].astype(np.int32)


def hash_inputs(n, n_volumes, log2T, seed=0, along_rays=True):
    rng = np.random.RandomState(seed)
    local = 1 << log2T
    feat = rng.uniform(-1e-2, 1e-2, size=(16 * local, 2)).astype(np.float32)
    prim = lcg_primes(16 * n_volumes * 3, seed + 1).reshape(16, n_volumes, 3)
    bias = np.zeros((16 * n_volumes, 3), np.float32)
    if along_rays:
        # samples in runs along rays, like the sampler emits them: (warp+1.5)/3 with warp in ~[-1,1]
        n_runs = max(1, n // 64)
        start = rng.uniform(0.2, 0.8, size=(n_runs, 1, 3))
        d = rng.normal(size=(n_runs, 1, 3))
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        k = np.arange(64).reshape(1, 64, 1)
        pts = (start + d * k / 768.0).reshape(-1, 3)[:n]
        if pts.shape[0] < n:
            pts = np.concatenate([pts, rng.uniform(0.17, 0.83, size=(n - pts.shape[0], 3))])
        anchors = np.repeat(rng.randint(0, n_volumes, size=n_runs), 64)[:n]
        if anchors.shape[0] < n:
            anchors = np.concatenate([anchors, rng.randint(0, n_volumes, size=n - anchors.shape[0])])
    else:
        pts = rng.uniform(0.17, 0.83, size=(n, 3))
        anchors = rng.randint(0, n_volumes, size=n)
    return feat, prim, bias, np.clip(pts, 0, 1).astype(np.float32), anchors.astype(np.int64)


def load_rig(name="rig8"):
    """Prebuilt synthetic aerial rig + octree blobs (tools/make_rig_fixture.py)."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    return {k: d[k] for k in d.files}


def rig_octree(rig):
    """PersOctree object around fixture blobs (no construction)."""
    from gfnerf_b200 import persoctree as po
    oc = po.PersOctree.__new__(po.PersOctree)
    oc.load_blobs(rig["tree_nodes"], rig["pers_trans"])
    n = oc.nodes.shape[0]
    oc.weight_stats = np.full(n, po.INIT_NODE_STAT, np.int64)
    oc.alpha_stats = np.full(n, po.INIT_NODE_STAT, np.int64)
    oc.visit_cnt = np.zeros(n, np.int64)
    oc.search_order = po.search_order_table()
    return oc


def make_sampler(rig, mode=1, device=None):
    from gfnerf_b200.perssampler import PersSamplerCore
    c2w = rig["c2w"]
    n = c2w.shape[0]
    w2c = np.tile(np.eye(4, dtype=np.float32)[None], (n, 1, 1))
    w2c[:, :3, :] = c2w
    w2c = np.linalg.inv(w2c)[:, :3, :].astype(np.float32)
    s = PersSamplerCore()
    s.InitSampler(1.5, [2000, 4000, 6000, 8000, 10000], 1000, 1024, 0.01, True, 10, 1.0 / 256, 16, c2w, w2c,
                  rig["intri"], rig["bounds"], mode, 512, 1.0, 16.0, 10000, device=device, octree=rig_octree(rig))
    return s
