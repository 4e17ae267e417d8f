"""Error-map importance sampling of training pixels for the focal (block) stage -- SURVEY.md 8(f) rank 4, the step
before ray generation -- and the error-map feedback after the forward pass.

Mirrors `ErrorPixelSampler` (nerfstudio/data/pixel_samplers.py:594-760, tensor-batch path) and
`TrainDataloader._update_error_map` (nerfstudio/data/utils/dataloaders.py:140-142) as called from
gfnerf/gf_pipeline.py:180-185: same method names, argument meaning and batch keys.

What is different, on purpose:
  * the batch (images, error map) is expected to be resident in HBM; nothing here moves it;
  * the 80 % uniform draw is done on the batch's device (`uniform_without_replacement`) instead of Python's
    `random.sample` over `range(n_pixels)` (6.5 k Python-level draws per step);
  * the `nonzero(error_map >= 0)` table the reference builds every step (n_pixels x 3 int64: 50 MB for 2 M pixels)
    is replaced by unravelling the flat pixel number, which is what that table holds whenever no entry is negative;
    with negative entries the reference's table is shorter than the flat weights it is indexed by -- that quirk is
    reproduced through the same `nonzero` call so behaviour (including the IndexError) is unchanged;
  * the error feedback is one kernel (`gf_error_map_update`, csrc/rays.cu) instead of abs / sum / index_put.

The weighted 20 % draw is `torch.multinomial` without replacement in chunks of 2^24, exactly as the reference:
which pixels come out is defined by torch's generator, so parity here is distributional (tests/test_pixel_samplers.py).
"""
from typing import Dict, Optional

import torch

from . import _lib

MULTINOMIAL_CHUNK = 2 ** 24      # torch.multinomial's category limit; pixel_samplers.py:647


def uniform_without_replacement(n: int, k: int, device, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """k distinct integers of [0, n), uniformly, in draw order (the reference: `random.sample(range(n), k)`).
    Draw k, re-draw the (rare) duplicates until none is left; expected rounds ~ 1 + k^2 / (2 n)."""
    if k < 0 or k > n:
        raise ValueError("Sample larger than population or is negative")        # random.sample's own error
    if k == 0:
        return torch.empty(0, dtype=torch.int64, device=device)
    if 2 * k > n:                                                              # dense case: a permutation is cheaper
        return torch.randperm(n, device=device, generator=generator)[:k]
    idx = torch.randint(n, (k,), device=device, generator=generator)
    while True:
        s, order = idx.sort(stable=True)
        dup_sorted = torch.zeros(k, dtype=torch.bool, device=device)
        dup_sorted[1:] = s[1:] == s[:-1]
        n_dup = int(dup_sorted.sum())
        if n_dup == 0:
            return idx
        idx[order[dup_sorted]] = torch.randint(n, (n_dup,), device=device, generator=generator)


class ErrorPixelSampler:
    """Samples pixel batches from an image batch: `weighted_choice_ratio` of them in proportion to the error map,
    the rest uniformly."""

    def __init__(self, num_rays_per_batch: int, keep_full_image: bool = False,
                 generator: Optional[torch.Generator] = None, **kwargs) -> None:
        self.kwargs = kwargs
        self.num_rays_per_batch = num_rays_per_batch
        self.keep_full_image = keep_full_image
        self.weighted_choice_ratio = 0.2                                       # pixel_samplers.py:606
        self.generator = generator

    def set_num_rays_per_batch(self, num_rays_per_batch: int):
        self.num_rays_per_batch = num_rays_per_batch

    def weighted_choice_multinomial(self, dist: torch.Tensor, size: int, device) -> torch.Tensor:
        """`size` flat pixel numbers drawn without replacement in proportion to `dist` (pixel_samplers.py:634-669):
        one multinomial if n < 2^24; else size // n_chunks from every full 2^24-chunk (each normalised on its own)
        and the size % n_chunks remainder from the tail chunk, or uniformly if n is a multiple of n_chunks."""
        n = dist.shape[0]
        n_chunks = n // MULTINOMIAL_CHUNK
        g = self.generator
        if n_chunks == 0:
            return torch.multinomial(dist, size, False, generator=g).to(device).long()
        picks = []
        for i in range(n_chunks):
            lo = i * MULTINOMIAL_CHUNK
            picks.append(torch.multinomial(dist[lo:lo + MULTINOMIAL_CHUNK], size // n_chunks, False, generator=g) + lo)
        rest = size % n_chunks
        if rest:
            if n % n_chunks != 0:
                lo = n_chunks * MULTINOMIAL_CHUNK
                picks.append(torch.multinomial(dist[lo:], rest, False, generator=g) + lo)
            else:
                picks.append(uniform_without_replacement(n, rest, dist.device, g))
        out = torch.cat(picks, dim=0).long().to(device)
        assert out.shape[0] == size
        return out

    def sample_method(self, batch_size: int, num_images: int, image_height: int, image_width: int,
                      error_map: torch.Tensor, device="cpu") -> torch.Tensor:
        """int64 [batch_size, 3] = (image slot, y, x): first int(batch_size * ratio) error-weighted rows, then the
        uniform ones (pixel_samplers.py:672-716)."""
        if error_map.dim() == 4:
            error_map = error_map.squeeze(-1)                                  # (n_image, h, w)
        weights = error_map.reshape(-1)
        n_pixels = weights.shape[0]
        n_weighted = int(batch_size * self.weighted_choice_ratio)
        chosen = torch.cat((self.weighted_choice_multinomial(weights, n_weighted, device),
                            uniform_without_replacement(n_pixels, batch_size - n_weighted, weights.device,
                                                        self.generator).to(device)), dim=0)
        if bool((weights >= 0).all()):
            h, w = error_map.shape[1], error_map.shape[2]
            indices = torch.stack((chosen // (h * w), (chosen // w) % h, chosen % w), dim=-1)
        else:                                                                  # the reference's table, quirk included
            indices = torch.nonzero(error_map >= 0.0, as_tuple=False).to(device)[chosen]
        assert indices.shape[0] == batch_size
        return indices.long()

    def collate_image_dataset_batch(self, batch: Dict, num_rays_per_batch: int, keep_full_image: bool = False):
        """Gathers every per-pixel entry of `batch` at the sampled pixels (pixel_samplers.py:718-757).  Returned keys:
        the batch's own per-pixel keys (`image`, `error_map`, masks ...), `indices` (column 0 = absolute camera
        index, `image_idx[slot]`), `rel_camera_indices`, and `full_image` if asked for."""
        device = batch["image"].device
        num_images, image_height, image_width, _ = batch["image"].shape
        indices = self.sample_method(num_rays_per_batch, num_images, image_height, image_width,
                                     error_map=batch["error_map"], device=device)
        c, y, x = indices[:, 0], indices[:, 1], indices[:, 2]
        out = {k: v[c, y, x] for k, v in batch.items() if k not in ("image_idx", "rel_camera_idx") and v is not None}
        assert out["image"].shape == (num_rays_per_batch, 3), out["image"].shape
        out["slot_indices"] = indices.clone()              # (slot, y, x): what update_error_map needs (see there)
        indices = indices.clone()
        indices[:, 0] = batch["image_idx"].to(device)[c]
        out["indices"] = indices
        out["rel_camera_indices"] = batch["rel_camera_idx"].to(device)[c]
        if keep_full_image:
            out["full_image"] = batch["image"]
        return out

    def sample(self, image_batch: Dict):
        if isinstance(image_batch["image"], torch.Tensor):
            return self.collate_image_dataset_batch(image_batch, self.num_rays_per_batch,
                                                    keep_full_image=self.keep_full_image)
        raise ValueError("image_batch['image'] must be a torch.Tensor (the list path of the reference, "
                         "pixel_samplers.py:759-829, calls sample_method with arguments it does not take)")


def update_error_map(error_map: torch.Tensor, ray_indices: torch.Tensor, pred_rgb: torch.Tensor,
                     gt_rgb: torch.Tensor, return_error: bool = False, check_indices: bool = True):
    """error_map[ray_indices[:,0], ray_indices[:,1], ray_indices[:,2]] = sum_c |gt - pred|, in place, one kernel.

    The reference passes `batch["indices"]`, whose column 0 is the ABSOLUTE camera index, and indexes the cached
    batch's error map (ordered by slot) with it (gf_pipeline.py:185, dataloaders.py:142): right only when the cached
    batch holds every image in dataset order, which is the configuration GF-NeRF trains in.  This function indexes
    with whatever it is given, like the reference; `ErrorPixelSampler` also returns `slot_indices` for callers whose
    cached batch is a subset.

    `check_indices` reads the kernel's out-of-range flag back (one 4-byte D2H, which waits for the forward that
    produced `pred_rgb`) to raise torch's IndexError; a training loop that trusts its sampler passes False and the
    bad rows are simply skipped."""
    _lib.require_cuda(error_map, ray_indices, pred_rgb, gt_rgb)
    if error_map.dtype != torch.float32 or not error_map.is_contiguous():
        raise RuntimeError("update_error_map: error_map must be a contiguous f32 tensor [n,h,w] or [n,h,w,1]")
    if error_map.dim() == 4 and error_map.shape[-1] == 1:
        n_img, h, w = error_map.shape[:3]
    elif error_map.dim() == 3:
        n_img, h, w = error_map.shape
    else:
        raise RuntimeError("update_error_map: error_map must be [n,h,w] or [n,h,w,1]")
    n = ray_indices.shape[0]
    if ray_indices.shape != (n, 3) or pred_rgb.shape != (n, 3) or gt_rgb.shape != (n, 3):
        raise RuntimeError("update_error_map: ray_indices i64 [n,3], pred_rgb / gt_rgb f32 [n,3]")
    idx = ray_indices.to(torch.int64).contiguous()
    pred, gt = pred_rgb.detach().float().contiguous(), gt_rgb.detach().float().contiguous()
    err = torch.empty(n, dtype=torch.float32, device=error_map.device) if return_error else None
    bad = torch.zeros(1, dtype=torch.int32, device=error_map.device)
    with torch.cuda.device(error_map.device):
        _lib.check(_lib.lib().gf_error_map_update(n, _lib.ptr(idx), _lib.ptr(pred), _lib.ptr(gt), n_img, h, w,
                                                  _lib.ptr(error_map), _lib.ptr(err), _lib.ptr(bad),
                                                  _lib.cur_stream()), "gf_error_map_update")
    if check_indices and int(bad.item()):
        raise IndexError("update_error_map: ray index out of range for the error map")   # torch raises here too
    return err
