"""Profiling aid (not part of the product): per-phase clock64() trace of one tile of the tcgen05 MLP backward.

  python tools/mlp_trace.py build     # here: nvcc -DGF_MLP_TRACE -> tools/trace/libgfnerf_b200_trace.so
  python tools/mlp_trace.py           # on the GPU box: run gf_mlp_backward on random data, print the phase deltas
"""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "trace", "libgfnerf_b200_trace.so")


def build():
    srcs = sorted(glob.glob(os.path.join(ROOT, "gf-nerf_b200", "csrc", "*.cu")))
    import importlib.util
    spec = importlib.util.spec_from_file_location("gf_build", os.path.join(ROOT, "gf-nerf_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = [mod.NVCC] + mod.FLAGS + ["-DGF_MLP_TRACE", "-shared", "-o", SO] + srcs
    subprocess.check_call(cmd)


def main():
    import numpy as np
    import torch
    L = C.CDLL(SO)
    n, R = 148 * 2 * 128 * 24, 1200      # ~760 samples per ray, like the bench rig
    g = torch.Generator(device="cuda").manual_seed(0)
    params = (torch.randn(11603, device="cuda", generator=g) * 0.1).contiguous()
    feat = (torch.randn(n, 32, device="cuda", generator=g) * 0.1).half().contiguous()
    ray_id = (torch.arange(n, device="cuda") // (n // R)).int().clamp_(max=R - 1).contiguous()
    ray_bias = (torch.randn(R, 64, device="cuda", generator=g) * 0.1).contiguous()
    d_sigma = torch.randn(n, device="cuda", generator=g).contiguous()
    d_rgb = torch.randn(n, 3, device="cuda", generator=g).contiguous()
    d_feat = torch.empty(n, 32, device="cuda", dtype=torch.float16)
    d_params = torch.zeros(11603, device="cuda")
    d_rb = torch.zeros(R, 64, device="cuda")
    vp, i64, f32, cint = C.c_void_p, C.c_int64, C.c_float, C.c_int
    L.gf_mlp_backward.argtypes = [i64, vp, cint, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp]
    masks = torch.full((n, 2, 4), -1, device="cuda", dtype=torch.int32)   # all units active
    st = torch.cuda.current_stream().cuda_stream
    for full in (1, 0):
        for _ in range(3):
            rc = L.gf_mlp_backward(n, None, 64, params.data_ptr(), feat.data_ptr(), ray_id.data_ptr(), ray_bias.data_ptr(),
                                   masks.data_ptr(), d_sigma.data_ptr(), d_rgb.data_ptr(), d_feat.data_ptr(),
                                   d_params.data_ptr() if full else None, d_rb.data_ptr() if full else None, 4096.0, st)
            assert rc == 0, L.gf_last_error()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.gf_mlp_backward(n, None, 64, params.data_ptr(), feat.data_ptr(), ray_id.data_ptr(), ray_bias.data_ptr(),
                          masks.data_ptr(), d_sigma.data_ptr(), d_rgb.data_ptr(), d_feat.data_ptr(),
                          d_params.data_ptr() if full else None, d_rb.data_ptr() if full else None, 4096.0, st)
        e1.record()
        torch.cuda.synchronize()
        print("mode", "full" if full else "frozen", "ms", e0.elapsed_time(e1), "tiles/CTA", n // 128 // 296,
              "ns/tile/CTA", e0.elapsed_time(e1) * 1e6 / (n // 128 // 296))
        buf = (C.c_ulonglong * 128)()
        assert L.gf_debug_mlp_trace(buf) == 0
        t = np.array(buf[:], dtype=np.int64).reshape(2, 64)
        names = ["start"]
        for r in range(5):
            names += [f"f{r}.E", f"f{r}.S", f"f{r}.I", f"f{r}.W"]
        for r in range(5):
            names += [f"b{r}.E", f"b{r}.S", f"b{r}.I", f"b{r}.W"] + ([f"b{r}.V"] if full else [])
        for th in range(2):
            print(" thread", th * 128)
            prev = t[th, 0]
            for i, nm in enumerate(names):
                if t[th, i] == 0:
                    continue
                print(f"   {nm:8s} +{t[th, i] - prev:6d}   (t={t[th, i] - t[th, 0]})")
                prev = t[th, i]


if __name__ == "__main__":
    build() if sys.argv[1:] == ["build"] else main()
