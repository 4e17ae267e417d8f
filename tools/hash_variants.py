"""Profiling aid (not part of the product): the two variants of the Hash3DAnchored forward gather (csrc/hash3d.cu:
PREFETCH = false, 48 registers / 5 CTAs per SM; PREFETCH = true, all 32 gathers of a four-level group in flight, 79
registers / 3 CTAs per SM) side by side on the samples the bench's sampler produces (rig20, 8192 rays), at log2T = 19
(table L2-resident), 21 and 23 (HBM).  One subprocess per variant (GF_HASH_FWD_VARIANT is read once per process).

  python tools/hash_variants.py            # on the GPU box
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker():
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from gfnerf_b200.hash_3d_anchored import Hash3DAnchoredCore
    from gfnerf_b200.persoctree import rig_rays
    from tests.helpers import make_sampler
    rig = bench.load_rig()
    s = make_sampler(rig, mode=1)
    o, d, _ = rig_rays(rig["c2w"], rig["intri"], 8192, seed=1234000)
    cs = s.sample_compact(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
    V = int(cs.total.item())
    pts, anc = cs.pts01[:V].contiguous(), cs.anchor[:V].contiguous()
    for log2T in (19, 21, 23):
        core = Hash3DAnchoredCore(log2T, s.n_volumes_)
        core.Reset()
        core.shadow(force=True)
        out = torch.empty((V, 32), dtype=torch.float16, device="cuda")
        fwd = lambda: core.launch_forward(pts, anc, out_f16=out, recast=False)
        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        ts = []
        for _ in range(9):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fwd()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(f"variant {os.environ.get('GF_HASH_FWD_VARIANT', 'auto'):>4s} log2T={log2T} {t:.3f} ms "
              f"{660 * V / t / 1e6:7.0f} GB/s algorithmic  checksum {float(out.double().sum()):.6f}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        worker()
    else:
        for v in ("0", "1", None):
            env = dict(os.environ)
            env.pop("GF_HASH_FWD_VARIANT", None)
            if v is not None:
                env["GF_HASH_FWD_VARIANT"] = v
            subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=env, check=True)
