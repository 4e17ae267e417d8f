"""Profiling aid (not part of the product): times build variants of the tcgen05 field-MLP kernels side by side.

  python tools/mlp_variants.py build     # here: one libgfnerf_b200 per variant under tools/trace/ (mlp_tc.cu recompiled
                                         # with the variant's -D flags, the other objects taken from gf-nerf_b200/build)
  python tools/mlp_variants.py           # on the GPU box: split forward / plain forward / full backward of each variant
                                         # on the bench's sample count, CUDA events, median of 7 launches
"""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tools", "trace")
VARIANTS = {
    "base": [],
}


def build():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gf_build", os.path.join(ROOT, "gf-nerf_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    objs = [o for o in sorted(glob.glob(os.path.join(ROOT, "gf-nerf_b200", "build", "*.o")))
            if os.path.basename(o) != "mlp_tc.o"]
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(ROOT, "gf-nerf_b200", "csrc", "mlp_tc.cu")
    for name, flags in VARIANTS.items():
        obj = os.path.join(OUT, f"mlp_tc_{name}.o")
        subprocess.check_call([mod.NVCC] + mod.FLAGS + flags + ["-c", src, "-o", obj])
        subprocess.check_call([mod.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fopenmp",
                               "-o", os.path.join(OUT, f"libvar_{name}.so")] + objs + [obj])
        print("built", name)


def main():
    import torch
    n, R = 6_600_000, 8192
    g = torch.Generator(device="cuda").manual_seed(0)
    params = (torch.randn(11603, device="cuda", generator=g) * 0.1).contiguous()
    feat = (torch.randn(n, 32, device="cuda", generator=g) * 0.3).half().contiguous()
    ray_id = (torch.arange(n, device="cuda") // (n // R + 1)).int().contiguous()
    ray_bias = (torch.randn(R, 64, device="cuda", generator=g) * 0.3).contiguous()
    d_sigma = (torch.randn(n, device="cuda", generator=g) * 1e-4).contiguous()
    d_rgb = (torch.randn(n, 3, device="cuda", generator=g) * 1e-4).contiguous()
    vp, i64, f32, cint = C.c_void_p, C.c_int64, C.c_float, C.c_int
    st = torch.cuda.current_stream().cuda_stream
    ref = None
    for name in VARIANTS:
        so = os.path.join(OUT, f"libvar_{name}.so")
        if not os.path.exists(so):
            continue
        L = C.CDLL(so)
        L.gf_last_error.restype = C.c_char_p
        L.gf_mlp_forward.argtypes = [i64, vp, cint, vp, vp, vp, vp, vp, vp, vp, vp]
        L.gf_mlp_backward.argtypes = [i64, vp, cint, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp]
        sigma, rgb = torch.empty(n, device="cuda"), torch.empty(n, 3, device="cuda")
        masks = torch.zeros(n, 2, 4, device="cuda", dtype=torch.int32)   # gf_mlp_mask_words(64) = 8
        d_feat = torch.empty(n, 32, device="cuda", dtype=torch.float16)
        d_params, d_rb = torch.zeros(11603, device="cuda"), torch.zeros(R, 64, device="cuda")

        def fwd(m):
            rc = L.gf_mlp_forward(n, None, 64, params.data_ptr(), feat.data_ptr(), ray_id.data_ptr(), ray_bias.data_ptr(),
                                  sigma.data_ptr(), rgb.data_ptr(), m, st)
            assert rc == 0, L.gf_last_error()

        def bwd():
            rc = L.gf_mlp_backward(n, None, 64, params.data_ptr(), feat.data_ptr(), ray_id.data_ptr(), ray_bias.data_ptr(),
                                   masks.data_ptr(), d_sigma.data_ptr(), d_rgb.data_ptr(), d_feat.data_ptr(),
                                   d_params.data_ptr(), d_rb.data_ptr(), 8192.0, st)
            assert rc == 0, L.gf_last_error()

        def timeit(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return sorted(ts)[len(ts) // 2]

        t_split = timeit(lambda: fwd(masks.data_ptr()))
        chk = (float(sigma.double().sum()), float(rgb.double().sum()), int(masks.long().sum()))
        t_plain = timeit(lambda: fwd(None))
        d_params.zero_()
        t_bwd = timeit(bwd)
        chk = chk + (float(d_feat.double().abs().sum()),)
        if ref is None:
            ref = chk
        print(f"{name:14s} fwd split {t_split:.3f} ms | fwd plain {t_plain:.3f} ms | bwd full {t_bwd:.3f} ms | "
              f"same results as base: {chk == ref}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        main()
