"""Scratch: per-step device time of the bench loop (one CUDA event per step) -- median vs mean, drift, outliers."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gfnerf_b200.engine import GFNeRFEngine
from tests.helpers import make_sampler
rig = bench.load_rig(); dev = torch.device("cuda", 0)
s = make_sampler(rig, mode=0, device=dev); s.ray_march_fineness_decay_end_iter_ = 0.0; s.ray_march_fineness_ = 1.0
s.generator = torch.Generator(device=dev).manual_seed(1234)
eng = GFNeRFEngine(s, log2_table_size=19, num_images=rig["c2w"].shape[0], seed=0)
host = bench.make_batches(rig, 8192, 8, seed=1234)
res = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in host]
N = int(os.environ.get("N", 400))
for i in range(10):
    o, d, cam, tgt = res[i % 8]; eng.train_step(o, d, tgt, cam, next_rays=res[(i + 1) % 8][:2])
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
ht = []
LOOK = int(os.environ.get("LOOKAHEAD", 0))     # > 0: the host never runs more than LOOK steps ahead of the device
ev[0].record()
for i in range(N):
    o, d, cam, tgt = res[(10 + i) % 8]
    if LOOK and i >= LOOK:
        ev[i + 1 - LOOK].synchronize()
    eng.train_step(o, d, tgt, cam, next_rays=res[(11 + i) % 8][:2])
    ev[i + 1].record()
    ht.append(time.perf_counter())
torch.cuda.synchronize()
dt = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(N)])
ht = np.diff(np.array(ht)) * 1e3
print("LOOKAHEAD", LOOK, "device ms/step: mean %.3f median %.3f p10 %.3f p90 %.3f max %.3f" % (dt.mean(), np.median(dt), np.percentile(dt, 10), np.percentile(dt, 90), dt.max()))
for a in range(0, N, 50):
    print("  steps %3d-%3d: device mean %.3f median %.3f | host issue mean %.3f max %.3f" % (a, a + 49, dt[a:a + 50].mean(), np.median(dt[a:a + 50]), ht[a:a + 49].mean(), ht[a:a + 49].max()))
