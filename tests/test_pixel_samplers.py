"""ErrorPixelSampler mirror (gfnerf_b200/pixel_samplers.py; reference nerfstudio/data/pixel_samplers.py:594-760) and
the error-map feedback kernel gf_error_map_update (reference gfnerf/gf_pipeline.py:180-185,
nerfstudio/data/utils/dataloaders.py:140-142).

Which pixels the sampler draws is defined by the random generators (torch.multinomial + random.sample in the
reference), so the host-logic tests check what the reference guarantees: shapes, ranges, distinctness, the 20 / 80
split, proportionality to the error map, the chunking rule above 2^24 pixels, and the collated batch's keys.  The
kernel is checked against the oracle (bit-exact: three fp32 subtractions and two additions in a fixed order)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from gfnerf_b200 import pixel_samplers as ps


def make_batch(n_img=5, h=24, w=32, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(n_img, h, w, 3, generator=g)
    err = torch.rand(n_img, h, w, 1, generator=g) * 0.01
    err[2, 4:12, 8:24] = 5.0                                   # a hot 8 x 16 region in image 2: ~98 % of the mass
    return {"image": image.to(device), "error_map": err.to(device),
            "image_idx": torch.tensor([10, 11, 12, 13, 14], device=device)[:n_img],
            "rel_camera_idx": torch.tensor([0, 1, 2, 3, 4], device=device)[:n_img],
            "mask": None}


def test_uniform_without_replacement_is_distinct_and_uniform():
    g = torch.Generator().manual_seed(1)
    idx = ps.uniform_without_replacement(5000, 2000, "cpu", g)           # many first-round duplicates
    assert idx.shape == (2000,) and idx.dtype == torch.int64
    assert idx.unique().numel() == 2000 and int(idx.min()) >= 0 and int(idx.max()) < 5000
    idx = ps.uniform_without_replacement(100, 90, "cpu", g)              # dense case
    assert idx.unique().numel() == 90
    assert ps.uniform_without_replacement(7, 0, "cpu", g).numel() == 0
    with pytest.raises(ValueError):                                      # random.sample's error (pixel_samplers.py:699)
        ps.uniform_without_replacement(5, 6, "cpu", g)
    counts = torch.zeros(50)
    for _ in range(400):
        counts[ps.uniform_without_replacement(50, 5, "cpu", g)] += 1
    assert float(counts.min()) > 15 and float(counts.max()) < 70         # expectation 40 each


def test_sample_method_split_and_proportionality():
    b = make_batch()
    s = ps.ErrorPixelSampler(1000, generator=torch.Generator().manual_seed(2))
    idx = s.sample_method(1000, 5, 24, 32, error_map=b["error_map"])
    assert idx.shape == (1000, 3) and idx.dtype == torch.int64
    assert int(idx[:, 0].max()) < 5 and int(idx[:, 1].max()) < 24 and int(idx[:, 2].max()) < 32 and int(idx.min()) >= 0
    hot = (idx[:, 0] == 2) & (idx[:, 1] >= 4) & (idx[:, 1] < 12) & (idx[:, 2] >= 8) & (idx[:, 2] < 24)
    # the first int(1000 * 0.2) rows are the error-weighted ones: 128 hot pixels hold 98 % of the mass and are drawn
    # without replacement, so (nearly) all 128 come out in those rows
    assert int(hot[:200].sum()) >= 120
    flat = (idx[:200, 0] * 24 + idx[:200, 1]) * 32 + idx[:200, 2]
    assert flat.unique().numel() == 200
    # the other 800 are uniform: the hot region is 128 / 3840 of the pixels
    assert int(hot[200:].sum()) < 80
    flat = (idx[200:, 0] * 24 + idx[200:, 1]) * 32 + idx[200:, 2]
    assert flat.unique().numel() == 800


def test_weighted_choice_chunking_rule(monkeypatch):
    """Above 2^24 pixels the reference splits the draw: size // n_chunks from every full chunk, the remainder from the
    tail.  Shrink the chunk to make the rule testable."""
    monkeypatch.setattr(ps, "MULTINOMIAL_CHUNK", 64)
    s = ps.ErrorPixelSampler(10, generator=torch.Generator().manual_seed(3))
    dist = torch.ones(64 * 3 + 10)
    out = s.weighted_choice_multinomial(dist, 31, "cpu")                 # 3 chunks x 10 + 1 from the tail
    assert out.shape == (31,)
    per_chunk = [int(((out >= 64 * i) & (out < 64 * (i + 1))).sum()) for i in range(3)]
    assert per_chunk == [10, 10, 10] and int((out >= 192).sum()) == 1
    dist = torch.ones(64 * 4)                                            # n % n_chunks == 0: remainder drawn uniformly
    out = s.weighted_choice_multinomial(dist, 9, "cpu")
    assert out.shape == (9,) and int(out.max()) < 256


def test_collate_keys_and_index_columns():
    b = make_batch()
    s = ps.ErrorPixelSampler(64, keep_full_image=True, generator=torch.Generator().manual_seed(4))
    out = s.sample(b)
    assert set(out) == {"image", "error_map", "indices", "slot_indices", "rel_camera_indices", "full_image"}
    assert out["image"].shape == (64, 3) and out["error_map"].shape == (64, 1)
    sl = out["slot_indices"]
    assert torch.equal(out["image"], b["image"][sl[:, 0], sl[:, 1], sl[:, 2]])
    assert torch.equal(out["indices"][:, 0], sl[:, 0] + 10)             # absolute camera index = image_idx[slot]
    assert torch.equal(out["indices"][:, 1:], sl[:, 1:])
    assert torch.equal(out["rel_camera_indices"], sl[:, 0])
    assert out["full_image"] is b["image"]
    s.set_num_rays_per_batch(32)
    assert s.sample(b)["image"].shape == (32, 3)
    with pytest.raises(ValueError):
        s.sample({"image": [b["image"][0]]})


def test_negative_entries_follow_the_reference_table():
    """With negative entries the reference indexes a SHORTER nonzero table with flat pixel numbers."""
    em = torch.ones(1, 4, 4, 1)
    em[0, 0, 0, 0] = -1.0
    s = ps.ErrorPixelSampler(4, generator=torch.Generator().manual_seed(5))
    table = torch.nonzero(em.squeeze(-1) >= 0.0, as_tuple=False)
    assert table.shape[0] == 15
    # multinomial rejects negative weights, as in the reference
    with pytest.raises(RuntimeError):
        s.sample_method(4, 1, 4, 4, error_map=em)


@pytest.mark.skipif(not os.path.isdir("/root/reference/nerfstudio"), reason="reference tree only in the build container")
def test_distribution_matches_the_reference_class():
    """The reference's own ErrorPixelSampler and the mirror on the same error map: per-image pick frequencies of the
    weighted and the uniform part agree within sampling noise."""
    tt = types.ModuleType("torchtyping")

    class TensorType:
        def __class_getitem__(cls, item):
            return cls

    tt.TensorType = TensorType
    tt.patch_typeguard = lambda: None
    sys.modules.setdefault("torchtyping", tt)
    spec = importlib.util.spec_from_file_location("ref_pixel_samplers",
                                                  "/root/reference/nerfstudio/data/pixel_samplers.py")
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference pixel_samplers.py does not import here: {e}")
    g = torch.Generator().manual_seed(7)
    em = torch.rand(4, 16, 16, 1, generator=g) * torch.tensor([1.0, 4.0, 0.25, 2.0]).view(4, 1, 1, 1)
    ref = mod.ErrorPixelSampler(500)
    mine = ps.ErrorPixelSampler(500, generator=g)
    fr, fm = torch.zeros(2, 4), torch.zeros(2, 4)
    for _ in range(60):
        a = ref.sample_method(500, 4, 16, 16, error_map=em)
        b = mine.sample_method(500, 4, 16, 16, error_map=em)
        assert a.shape == b.shape and a.dtype == b.dtype
        for part, sl in enumerate((slice(0, 100), slice(100, 500))):
            fr[part] += torch.bincount(a[sl, 0], minlength=4)
            fm[part] += torch.bincount(b[sl, 0], minlength=4)
    fr, fm = fr / fr.sum(1, keepdim=True), fm / fm.sum(1, keepdim=True)
    assert float((fr - fm).abs().max()) < 0.02, (fr, fm)
    assert float((fm[1] - 0.25).abs().max()) < 0.02                      # the uniform part


def test_update_error_map_refuses_host_tensors():
    with pytest.raises(RuntimeError, match="no CPU path"):
        ps.update_error_map(torch.zeros(1, 2, 2), torch.zeros(1, 3, dtype=torch.int64), torch.zeros(1, 3),
                            torch.zeros(1, 3))


@pytest.mark.gpu
def test_error_map_update_matches_oracle():
    from oracle import oracle as orc
    rng = np.random.RandomState(0)
    n_img, h, w, n = 6, 40, 56, 4096
    em = rng.uniform(0, 1, size=(n_img, h, w)).astype(np.float32)
    flat = rng.choice(n_img * h * w, size=n, replace=False)              # distinct pixels: a defined result
    idx = np.stack(np.unravel_index(flat, (n_img, h, w)), axis=-1).astype(np.int64)
    idx[:5, 0] -= n_img                                                  # negative indices wrap once, as in torch
    pred = rng.uniform(0, 1, size=(n, 3)).astype(np.float32)
    gt = rng.uniform(0, 1, size=(n, 3)).astype(np.float32)
    want = em.copy()
    want_err = orc.error_map_update(want, idx, pred, gt)
    dev = torch.device("cuda:0")
    em_d = torch.tensor(em, device=dev).unsqueeze(-1)                    # [n,h,w,1] as the reference holds it
    err = ps.update_error_map(em_d, torch.tensor(idx, device=dev), torch.tensor(pred, device=dev),
                              torch.tensor(gt, device=dev), return_error=True)
    assert np.array_equal(err.cpu().numpy(), want_err)
    assert np.array_equal(em_d[..., 0].cpu().numpy(), want)
    # empty batch is a no-op; an out-of-range row raises like torch's index_put and leaves the map alone
    ps.update_error_map(em_d, torch.zeros(0, 3, dtype=torch.int64, device=dev), torch.zeros(0, 3, device=dev),
                        torch.zeros(0, 3, device=dev))
    bad = torch.tensor([[n_img, 0, 0]], dtype=torch.int64, device=dev)
    with pytest.raises(IndexError):
        ps.update_error_map(em_d, bad, torch.zeros(1, 3, device=dev), torch.ones(1, 3, device=dev))
    assert np.array_equal(em_d[..., 0].cpu().numpy(), want)


@pytest.mark.gpu
def test_sampler_on_device_batch():
    dev = torch.device("cuda:0")
    b = make_batch(device=dev)
    s = ps.ErrorPixelSampler(256, generator=torch.Generator(device=dev).manual_seed(1))
    out = s.sample(b)
    assert out["image"].is_cuda and out["indices"].is_cuda and out["image"].shape == (256, 3)
    sl = out["slot_indices"]
    assert torch.equal(out["image"], b["image"][sl[:, 0], sl[:, 1], sl[:, 2]])
    ps.update_error_map(b["error_map"], sl, out["image"] * 0.5, out["image"])
    got = b["error_map"][sl[:, 0], sl[:, 1], sl[:, 2], 0]
    assert torch.allclose(got, (out["image"] * 0.5).sum(-1), rtol=1e-6, atol=1e-7)
