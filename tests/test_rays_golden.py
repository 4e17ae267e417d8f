"""Ray generation: the oracle and the CUDA kernel against vectors produced by RUNNING the
reference's `Cameras.generate_rays` (tests/golden/make_golden_rays.py -> ref_rays.npz)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_rays.npz")


def _check(got, d):
    # origins / lookat are copies; directions and their norm follow torch's op sequence bit for bit
    for k in ("origins", "directions", "lookat", "dir_norm"):
        assert np.array_equal(np.asarray(got[k]).reshape(d[k].shape), d[k]), k
    # pixel_area = dx * dy of differences of nearly equal unit vectors: torch.sum's vectorised reduction rounds the
    # last bit differently in < 1 % of the rays
    pa = np.asarray(got["pixel_area"]).reshape(-1)
    assert np.allclose(pa, d["pixel_area"], rtol=3e-7, atol=0)
    assert (pa != d["pixel_area"]).mean() < 0.02


def test_oracle_matches_the_reference_output():
    d = np.load(GOLD)
    _check(orc.generate_rays(d["cam_idx"], d["coords"], d["c2w"], d["fx"], d["fy"], d["cx"], d["cy"]), d)


@pytest.mark.gpu
def test_kernel_matches_the_reference_output_and_the_oracle():
    import torch
    from gfnerf_b200 import Cameras
    d = np.load(GOLD)
    T = lambda a: torch.from_numpy(a).cuda()
    cams = Cameras(T(d["c2w"]), T(d["fx"]), T(d["fy"]), T(d["cx"]), T(d["cy"]), width=1920, height=1080)
    rb = cams.generate_rays(T(d["cam_idx"])[:, None], T(d["coords"]))
    got = {"origins": rb.origins.cpu().numpy(), "directions": rb.directions.cpu().numpy(),
           "lookat": rb.lookat_directions.cpu().numpy(), "pixel_area": rb.pixel_area.cpu().numpy(),
           "dir_norm": rb.metadata["directions_norm"].cpu().numpy()}
    _check(got, d)
    ref = orc.generate_rays(d["cam_idx"], d["coords"], d["c2w"], d["fx"], d["fy"], d["cx"], d["cy"])
    for k in ("origins", "directions", "lookat", "dir_norm", "pixel_area"):
        assert np.array_equal(got[k].reshape(ref[k].shape), ref[k]), k           # kernel == oracle, bit for bit
    # a whole frame: unit directions, origins = the camera centre, pixel (0.5, 0.5) first
    frame = cams.generate_frame_rays(2)
    assert frame.origins.shape == (1920 * 1080, 3)
    assert float((frame.directions.norm(dim=-1) - 1).abs().max()) < 1e-6
    assert torch.equal(frame.origins[0], T(d["c2w"])[2, :, 3])
    one = cams.generate_rays(torch.tensor([2], device="cuda"), torch.tensor([[0.5, 0.5]], device="cuda"))
    assert torch.equal(one.directions[0], frame.directions[0])
    # errors / edge cases
    with pytest.raises(RuntimeError):
        cams.generate_rays(torch.tensor([99], device="cuda"), torch.tensor([[0.5, 0.5]], device="cuda"))
    with pytest.raises(RuntimeError):
        cams.generate_rays(torch.tensor([0]), torch.tensor([[0.5, 0.5]]))
    empty = cams.generate_rays(torch.zeros(0, dtype=torch.int64, device="cuda"), torch.zeros(0, 2, device="cuda"))
    assert empty.origins.shape == (0, 3)
