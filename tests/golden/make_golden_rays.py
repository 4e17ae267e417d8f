#!/usr/bin/env python
"""Generates tests/golden/ref_rays.npz by RUNNING THE REFERENCE's own `Cameras.generate_rays`
(nerfstudio/cameras/cameras.py:446-727, with GF-NeRF's `lookat_directions` addition at :704,723) on CPU: perspective
cameras without distortion, the configuration GF-NeRF's datamanager uses.

Run in the build container (needs /root/reference; it does not exist on the GPU box, so the vectors are committed):

  python tests/golden/make_golden_rays.py
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get("GF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    tt = types.ModuleType("torchtyping")

    class TensorType:
        def __class_getitem__(cls, item):
            return cls

    tt.TensorType = TensorType
    tt.patch_typeguard = lambda: None
    sys.modules["torchtyping"] = tt
    sys.modules["nerfacc"] = types.ModuleType("nerfacc")
    sys.path.insert(0, REF)
    from nerfstudio.cameras.cameras import Cameras, CameraType
    from gfnerf_b200.persoctree import aerial_rig
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    rng = np.random.RandomState(7)
    c2w, intri, _ = aerial_rig(n_side=4, extent=4.0, seed=3)
    n_cams = c2w.shape[0]
    # per-camera intrinsics (the rig's are shared; vary them so that the per-camera gather is exercised)
    fx = (intri[:, 0, 0] * rng.uniform(0.8, 1.2, n_cams)).astype(np.float32)
    fy = (intri[:, 1, 1] * rng.uniform(0.8, 1.2, n_cams)).astype(np.float32)
    cx = (intri[:, 0, 2] + rng.uniform(-20, 20, n_cams)).astype(np.float32)
    cy = (intri[:, 1, 2] + rng.uniform(-20, 20, n_cams)).astype(np.float32)
    W, H = 1920, 1080
    cams = Cameras(camera_to_worlds=torch.from_numpy(c2w), fx=torch.from_numpy(fx), fy=torch.from_numpy(fy),
                   cx=torch.from_numpy(cx), cy=torch.from_numpy(cy), width=W, height=H,
                   camera_type=CameraType.PERSPECTIVE)
    n = 4000
    cam_idx = rng.randint(0, n_cams, size=n).astype(np.int64)
    # pixel centres, as the pixel samplers produce them (image_coords = index + 0.5), corners and borders included
    yy = rng.randint(0, H, size=n).astype(np.float32) + 0.5
    xx = rng.randint(0, W, size=n).astype(np.float32) + 0.5
    yy[:4], xx[:4] = [0.5, 0.5, H - 0.5, H - 0.5], [0.5, W - 0.5, 0.5, W - 0.5]
    coords = np.stack([yy, xx], -1).astype(np.float32)
    rb = cams.generate_rays(camera_indices=torch.from_numpy(cam_idx)[:, None], coords=torch.from_numpy(coords))
    out = os.path.join(HERE, "ref_rays.npz")
    np.savez_compressed(out, c2w=c2w, fx=fx, fy=fy, cx=cx, cy=cy, cam_idx=cam_idx, coords=coords,
                        origins=rb.origins.numpy(), directions=rb.directions.numpy(),
                        lookat=rb.lookat_directions.numpy(), pixel_area=rb.pixel_area.numpy()[:, 0],
                        dir_norm=rb.metadata["directions_norm"].numpy()[:, 0])
    print("wrote", out, rb.origins.shape, float(rb.pixel_area.mean()))


if __name__ == "__main__":
    main()
