"""Ray generation for the per-ray path: the perspective / no-distortion case of the reference's
`Cameras.generate_rays` (nerfstudio/cameras/cameras.py:446-727, with GF-NeRF's `lookat_directions`, :704,723) as ONE
kernel (`gf_generate_rays`, csrc/rays.cu) instead of ~40 torch ops.  Same argument meaning as the reference:
`camera_indices` [n] or [n,1], `coords` [n,2] = pixel (y, x) as produced by the pixel samplers (index + 0.5)."""
from typing import Optional

import torch

from . import _lib
from .rays import RayBundle


class Cameras:
    """The subset of nerfstudio's `Cameras` the path needs: camera_to_worlds [n,3,4] and per-camera fx, fy, cx, cy."""

    def __init__(self, camera_to_worlds: torch.Tensor, fx, fy, cx, cy, width: Optional[int] = None,
                 height: Optional[int] = None):
        if not camera_to_worlds.is_cuda:
            raise RuntimeError("gfnerf_b200: tensor is not on a CUDA device (there is no CPU path)")
        dev = camera_to_worlds.device
        n = camera_to_worlds.shape[0]
        if camera_to_worlds.shape[1:] != (3, 4):
            raise RuntimeError("camera_to_worlds: f32 [n,3,4]")
        self.camera_to_worlds = camera_to_worlds.float().contiguous()

        def per_cam(v):
            t = torch.as_tensor(v, dtype=torch.float32, device=dev).reshape(-1)
            return (t.expand(n) if t.numel() == 1 else t).contiguous()

        self.fx, self.fy, self.cx, self.cy = per_cam(fx), per_cam(fy), per_cam(cx), per_cam(cy)
        self.width, self.height = width, height
        self.device = dev

    def __len__(self):
        return int(self.camera_to_worlds.shape[0])

    def get_intrinsics_matrices(self) -> torch.Tensor:
        """K [n,3,3] (nerfstudio/cameras/cameras.py get_intrinsics_matrices)"""
        K = torch.zeros((len(self), 3, 3), dtype=torch.float32, device=self.device)
        K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = self.fx, self.fy, self.cx, self.cy, 1.0
        return K

    @torch.no_grad()
    def generate_rays(self, camera_indices: torch.Tensor, coords: torch.Tensor) -> RayBundle:
        _lib.require_cuda(camera_indices, coords)
        idx = camera_indices.reshape(-1).to(torch.int64).contiguous()
        n = idx.shape[0]
        coords = coords.reshape(n, 2).float().contiguous()
        if n and (int(idx.min()) < 0 or int(idx.max()) >= len(self)):
            raise RuntimeError("generate_rays: camera index out of range")      # torch indexing raises too
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=self.device)
        o, d, la, pa, dn = f(n, 3), f(n, 3), f(n, 3), f(n, 1), f(n, 1)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gf_generate_rays(
                n, _lib.ptr(idx), _lib.ptr(coords), _lib.ptr(self.camera_to_worlds), _lib.ptr(self.fx),
                _lib.ptr(self.fy), _lib.ptr(self.cx), _lib.ptr(self.cy), len(self), _lib.ptr(o), _lib.ptr(d),
                _lib.ptr(la), _lib.ptr(pa), _lib.ptr(dn), _lib.cur_stream()), "gf_generate_rays")
        return RayBundle(origins=o, directions=d, lookat_directions=la, pixel_area=pa,
                         camera_indices=idx.view(n, 1), metadata={"directions_norm": dn})

    def generate_frame_rays(self, camera_index: int) -> RayBundle:
        """All pixels of one camera, row-major (the render loop's ray bundle; image_coords = index + 0.5)."""
        ys = torch.arange(self.height, device=self.device, dtype=torch.float32) + 0.5
        xs = torch.arange(self.width, device=self.device, dtype=torch.float32) + 0.5
        coords = torch.stack(torch.meshgrid(ys, xs, indexing="ij"), -1).reshape(-1, 2)
        idx = torch.full((coords.shape[0],), int(camera_index), dtype=torch.int64, device=self.device)
        return self.generate_rays(idx, coords)
