"""Where the operator-API step (GFNeRFModel.get_outputs + autograd + torch.optim.Adam, bench.py operator_api_arm) spends
its time: per-section wall time with a synchronize after each section, and the kernel-level top list of torch's profiler.
Scratch measurement tool (gpurun)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gfnerf_b200 as gf  # noqa: E402
from tests.helpers import rig_octree  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    rig = bench.load_rig()
    R = 8192
    ps = gf.PersSampler(rig["c2w"], rig["intri"], rig["bounds"], bbox_levels=10, mode=0, octree=rig_octree(rig),
                        ray_march_fineness_decay_end_iter=0, device=dev, seed=1234)
    field = gf.GFNeRFField(torch.zeros(2, 3), rig["c2w"].shape[0], log2_hashmap_size=19, use_appearance_embedding=True,
                           n_volumes=ps.get_n_volumes(), generator=torch.Generator().manual_seed(0)).to(dev)
    model = gf.GFNeRFModel(ps, field).to(dev)
    model.train()
    params = list(field.base_encoding_init.get_params()) + list(field.base_network.parameters()) + \
        list(field.mlp_head.parameters()) + list(field.embedding_appearance.parameters())
    opt = torch.optim.Adam(params, lr=1e-2, eps=1e-15)
    host = bench.make_batches(rig, R, 4, seed=4321)
    dev_b = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in host]
    ones = torch.ones(R, 1, device=dev)
    sec = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        sec[name] = sec.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def step(i, timed):
        o, d, cam, tgt = dev_b[i % 4]
        rb = gf.RayBundle(origins=o, directions=d, lookat_directions=d, pixel_area=ones, camera_indices=cam.view(-1, 1),
                          rel_camera_indices=cam.view(-1, 1), steps=torch.full((R, 1), 20001 + i, device=dev))
        t = time.perf_counter()
        rs = model.persampler(rb)
        if timed: t = tick("sampler module", t)
        fo = model.field(rs)
        if timed: t = tick("field", t)
        w, a, tr = rs.get_weights_f2nerf(fo[gf.FieldHeadNames.DENSITY])
        rgb = model.renderer_rgb(rgb=fo[gf.FieldHeadNames.RGB], weights=w)
        depth = model.renderer_depth(weights=w, ray_samples=rs)
        acc = model.renderer_accumulation(weights=w)
        if timed: t = tick("weights + renderers", t)
        model.persampler.update_oct_nodes(sampled_anchors=rs.f2samples.sampled_anchors, pts_idx_bounds=rs.f2samples.pts_idx_start_end,
                                          sampled_weights=w.detach(), sampled_alpha=a, iter_step=20001 + i)
        if timed: t = tick("update_oct_nodes", t)
        diff = rgb - tgt
        loss = torch.sqrt(diff * diff + 1e-12).sum() / R
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if timed: t = tick("loss + backward", t)
        opt.step()
        if timed: t = tick("optimizer", t)
        return loss

    for i in range(3):
        step(i, False)
    torch.cuda.synchronize()
    n = 5
    for i in range(n):
        step(3 + i, True)
    for k, v in sec.items():
        print(f"{k:24s} {v / n * 1e3:8.3f} ms")
    print(f"{'sum':24s} {sum(sec.values()) / n * 1e3:8.3f} ms")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        step(10 + i, False)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"unsynchronised: host issue {1e3 * (t1 - t0) / n:.3f} ms/step, total {1e3 * (t2 - t0) / n:.3f} ms/step")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(3):
            step(20 + i, False)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))


if __name__ == "__main__":
    main()
