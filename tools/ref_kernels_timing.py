"""The reference's OWN kernels (oracle/_ref/libgf_ref_cuda.so: the device functions of Hash3DAnchored_cuda.cu and
PersSampler_cuda.cu extracted at build time and compiled by nvcc for sm_100a, oracle/ref_driver_cuda.cu) timed on the
same B200, on the bench workload (BASELINE.json configs[1]: 400-camera rig, 8192 rays, log2T = 19), beside ours on the
same inputs.  Measurement infrastructure, not the product and not bench.py: run through gpurun, output committed under
profiles/.  The only "reference on the same box" number there is for SURVEY 8 rows a1 / a2 / a6 / a8 -- the reference's
full extension cannot be built (un-vendored tiny-cuda-nn + patched Eigen).

  python tools/ref_kernels_timing.py [--rays 8192] [--iters 5] > gpurun_out/rXX_ref_kernels_timing.json

What is timed (CUDA events on the default stream, median of --iters after one warm-up):
  * GetSamples: the reference's launch sequence (count pass, host sync, fill pass, march-count pass, host cumsum,
    march-fill pass: PersSampler_cuda.cu:321-477) vs our fused sample_rays + scan + compact;
  * Hash3DAnchored forward / backward: the reference's kernels on (a) the dense [R x 1024] slots the reference's field
    queries (nerfacto_field.py:437-455: padding included) and (b) only the V valid samples, vs ours on the V samples.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def med_ms(fn, iters):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return float(np.median(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--log2T", type=int, default=19)
    args = ap.parse_args()
    from oracle import oracle as orc
    from oracle import ref_host as rh
    from gfnerf_b200 import _lib
    from gfnerf_b200.engine import GFNeRFEngine
    from gfnerf_b200.persoctree import rig_rays
    from tests.helpers import load_rig, make_sampler
    assert rh.cuda_available(), "oracle/_ref/libgf_ref_cuda.so missing: run __graft_entry__.build() in the build container"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    rig = load_rig("rig20")
    R, S = args.rays, 1024
    sampler = make_sampler(rig, mode=1, device=dev)      # eval mode: noise = 1, the same samples on both sides
    eng = GFNeRFEngine(sampler, log2_table_size=args.log2T, num_images=rig["c2w"].shape[0], seed=0)
    o, d, _ = rig_rays(rig["c2w"], rig["intri"], R, seed=1234000)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    to, td = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    noise = torch.ones(S + R + 10, device=dev)
    so = torch.from_numpy(orc.search_order()).to(dev)
    res = {"workload": f"rig20 (400 cameras), {R} rays, log2T={args.log2T}, eval-mode march (noise 1)", "iters": args.iters}

    # ---- sampling ----------------------------------------------------------------------------------------------
    ref_s = rh.cuda_get_samples(to, td, noise, sampler.tree_nodes_gpu_, sampler.pers_trans_gpu_, so)
    cs = sampler.sample_compact(to, td)
    V = int(cs.total.item())
    assert torch.equal(cs.counts, ref_s["counts"]), "sample counts differ from the reference kernels"
    res["valid_samples"] = V
    res["get_samples_ms"] = {
        "reference_kernels": med_ms(lambda: rh.cuda_get_samples(to, td, noise, sampler.tree_nodes_gpu_,
                                                                sampler.pers_trans_gpu_, so), args.iters),
        "ours": med_ms(lambda: sampler.sample_compact(to, td), args.iters)}
    res["get_samples_ms"]["note"] = ("reference: its launch sequence incl. the two host round trips and the zero-fill "
                                     "of its dense [R,1024] outputs; ours: sample_rays + scan + compact, no host sync")

    # ---- hash encode -------------------------------------------------------------------------------------------
    enc = eng.enc
    table = enc.feat_pool_.detach()
    prim, bias = enc.prim_pool_, enc.bias_pool_
    m = torch.arange(S, device=dev)[None, :] < ref_s["counts"][:, None]
    pts_dense = ((ref_s["warp_pts"] + 1.5) / 3.0).reshape(-1, 3).contiguous()       # nerfacto_field.py:440
    anc_dense = ref_s["anchors"][..., 0].reshape(-1).contiguous()
    pts_v, anc_v = pts_dense[m.reshape(-1)].contiguous(), anc_dense[m.reshape(-1)].contiguous()
    assert pts_v.shape[0] == V
    L = rh.cuda_lib()
    import ctypes as C

    def ref_fwd(pts, anc, out):
        n = pts.shape[0]
        rc = L.refcu_hash_forward(C.c_int(n), C.c_int(prim.shape[1]), C.c_void_p(t16.data_ptr()), C.c_void_p(prim32.data_ptr()),
                                  C.c_void_p(fidx.data_ptr()), C.c_void_p(fsize.data_ptr()), C.c_void_p(bias.data_ptr()),
                                  C.c_void_p(pts.data_ptr()), C.c_void_p(anc.data_ptr()), C.c_void_p(out.data_ptr()))
        assert rc == 0

    def ref_bwd(pts, anc, gin, gout):
        n = pts.shape[0]
        gout.zero_()
        rc = L.refcu_hash_backward(C.c_int(n), C.c_int(prim.shape[1]), C.c_void_p(prim32.data_ptr()), C.c_void_p(fidx.data_ptr()),
                                   C.c_void_p(fsize.data_ptr()), C.c_void_p(bias.data_ptr()), C.c_void_p(pts.data_ptr()),
                                   C.c_void_p(anc.data_ptr()), C.c_void_p(gin.data_ptr()), C.c_void_p(gout.data_ptr()))
        assert rc == 0

    local = enc.local_size_
    t16 = table.half().contiguous()
    prim32 = prim.int().contiguous()
    fidx = (torch.arange(16, device=dev) * local).to(torch.int32)
    fsize = torch.full((16,), local, dtype=torch.int32, device=dev)
    out_dense = torch.zeros((R * S, 32), dtype=torch.float16, device=dev)
    out_v = torch.zeros((V, 32), dtype=torch.float16, device=dev)
    ours_v = torch.empty((V, 32), dtype=torch.float16, device=dev)
    anc_v32 = anc_v.to(torch.int32).contiguous()
    enc.shadow(force=True)
    ref_fwd(pts_v, anc_v, out_v)
    enc.launch_forward(pts_v, anc_v32, out_f16=ours_v, recast=False)
    torch.cuda.synchronize()
    res["hash_forward_bit_equal"] = bool(torch.equal(out_v, ours_v))
    res["hash_forward_ms"] = {
        "reference_kernel_dense_slots": med_ms(lambda: ref_fwd(pts_dense, anc_dense, out_dense), args.iters),
        "reference_kernel_valid_samples": med_ms(lambda: ref_fwd(pts_v, anc_v, out_v), args.iters),
        "reference_table_cast": med_ms(lambda: table.half(), args.iters),      # Hash3DAnchored_cuda.cu:185, every forward
        "ours_valid_samples": med_ms(lambda: enc.launch_forward(pts_v, anc_v32, out_f16=ours_v, recast=False), args.iters)}
    gin_dense = (torch.randn((R * S, 32), device=dev) * 1e-2).half()
    gin_dense[~m.reshape(-1)] = 0
    gin_v = gin_dense[m.reshape(-1)].contiguous()
    g16 = torch.zeros((16 * local, 2), dtype=torch.float16, device=dev)
    g32 = torch.zeros((16 * local, 2), dtype=torch.float32, device=dev)
    res["hash_backward_ms"] = {
        "reference_kernel_dense_slots": med_ms(lambda: ref_bwd(pts_dense, anc_dense, gin_dense, g16), args.iters),
        "reference_kernel_valid_samples": med_ms(lambda: ref_bwd(pts_v, anc_v, gin_v, g16), args.iters),
        "ours_valid_samples": med_ms(lambda: (g32.zero_(), enc.launch_backward(pts_v, anc_v32, gin_v, True, g32)),
                                     args.iters),
        "note": "gradient rows already x128 fp16 on both sides; both include the zero-fill of the gradient table"}
    for k in ("get_samples_ms", "hash_forward_ms", "hash_backward_ms"):
        r = res[k]
        ref_key = "reference_kernels" if "reference_kernels" in r else "reference_kernel_dense_slots"
        ours_key = "ours" if "ours" in r else "ours_valid_samples"
        r["speedup_vs_reference_as_called"] = round(r[ref_key] / r[ours_key], 2)
        if "reference_kernel_valid_samples" in r:
            r["speedup_same_points"] = round(r["reference_kernel_valid_samples"] / r[ours_key], 2)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
