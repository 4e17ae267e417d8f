"""GPU: the nerfstudio-side operator API (PersSampler module -> GFNeRFField -> RaySamples.get_weights_f2nerf ->
renderers, i.e. GFNeRFModel.get_outputs of gfnerf/nerfacto.py:522-619 on dense [R,1024] tensors with torch autograd)
against the fused engine on the same rays and parameters; the focal (block) stage; state-dict compatibility."""
import numpy as np
import pytest
import torch

from tests.helpers import load_rig, make_sampler, rig_octree

pytestmark = pytest.mark.gpu


def build(rig, log2T=14, seed=0, hidden=64):
    import gfnerf_b200 as gf
    from gfnerf_b200.engine import GFNeRFEngine
    eng = GFNeRFEngine(make_sampler(rig, mode=1), log2_table_size=log2T, num_images=rig["c2w"].shape[0], seed=seed,
                       hidden=hidden)
    eng.enc.feat_pool_.data.uniform_(-0.5, 0.5)
    eng.enc.shadow(force=True)
    ps = gf.PersSampler(rig["c2w"], rig["intri"], rig["bounds"], bbox_levels=10, mode=1, octree=rig_octree(rig))
    field = gf.GFNeRFField(torch.zeros(2, 3), rig["c2w"].shape[0], log2_hashmap_size=log2T, hidden_dim=hidden,
                           hidden_dim_color=hidden, use_appearance_embedding=True, n_volumes=eng.n_volumes).cuda()
    # identical parameters: table, primes, MLP blob, embedding
    field.base_encoding_init.load_states(eng.enc.States(), 0)
    o = 0
    for lin in field.base_network.linears() + field.mlp_head.linears():
        for p in (lin.weight, lin.bias):
            p.data.copy_(eng.mlp[o:o + p.numel()].view_as(p))
            o += p.numel()
    assert o == eng.mlp.numel()
    field.embedding_appearance.embedding.weight.data.copy_(eng.emb)
    model = gf.GFNeRFModel(ps, field).cuda()
    return eng, model


def bundle(rig, R, seed):
    import gfnerf_b200 as gf
    from gfnerf_b200.persoctree import rig_rays
    o, d, cam = rig_rays(rig["c2w"], rig["intri"], R, seed=seed)
    T = lambda a: torch.from_numpy(a).cuda()
    rb = gf.RayBundle(origins=T(o), directions=T(d), lookat_directions=T(d), pixel_area=torch.ones(R, 1).cuda(),
                      camera_indices=T(cam).view(-1, 1), rel_camera_indices=T(cam).view(-1, 1),
                      steps=torch.zeros(R, 1).cuda())
    return rb, T(o), T(d), T(cam)


@pytest.mark.parametrize("hidden", [64, 128])     # 128: the reference's shipped gf-nerf config (gfnerf/config.py:124-125)
def test_operator_path_equals_fused_engine(hidden):
    rig = load_rig("rig8")
    eng, model = build(rig, hidden=hidden)
    R = 384
    rb, o, d, cam = bundle(rig, R, seed=5)
    target = torch.rand(R, 3, generator=torch.Generator().manual_seed(1)).cuda()
    # fused engine: forward + backward, no optimizer step
    out = eng.train_step(o, d, target, cam, optimizer_step=False, update_octree=False)
    # operator API with autograd
    model.eval()                       # no octree feedback; renderers differ only by the eval clamp
    res = model.get_outputs(rb)
    assert res["rgb"].shape == (R, 3) and res["depth"].shape == (R, 1) and res["accumulation"].shape == (R, 1)
    assert torch.allclose(res["rgb"], out.rgb.clamp(0, 1), rtol=1e-4, atol=1e-5)
    assert torch.allclose(res["accumulation"][:, 0], out.accumulation, rtol=1e-4, atol=1e-5)
    assert torch.allclose(res["depth"][:, 0], out.depth, rtol=1e-4, atol=1e-5)
    assert float(res["oct_depth"].min()) > 0
    model.train()
    model.persampler.sampler.UpdateMode(1)
    rb.steps = None                    # keeps the octree untouched in train mode
    res = model.get_outputs(rb)
    diff = res["rgb"] - target
    loss = torch.sqrt(diff * diff + 1e-12).sum() / R          # CharbonnierLoss, losses.py:73-84
    loss.backward()
    assert abs(float(loss) - float(out.loss)) < 1e-4 * float(out.loss)
    f = model.field
    g_table = f.base_encoding_init.hash_3d.feat_pool_.grad
    ref = eng.opt_table.unscaled_grad().view_as(g_table)
    assert float((g_table - ref).abs().max()) <= 2e-3 * float(ref.abs().max())
    got = torch.cat([torch.cat([l.weight.grad.reshape(-1), l.bias.grad.reshape(-1)])
                     for l in f.base_network.linears() + f.mlp_head.linears()])
    assert float((got - eng.opt_mlp.grad).abs().max()) <= 2e-3 * float(eng.opt_mlp.grad.abs().max())
    ge = f.embedding_appearance.embedding.weight.grad
    assert float((ge - eng.opt_emb.grad.view_as(ge)).abs().max()) <= 2e-3 * float(eng.opt_emb.grad.abs().max())


def test_block_stage_trains_only_the_residual_table():
    rig = load_rig("rig8")
    _, model = build(rig)
    f = model.field
    R = 256
    rb, *_ = bundle(rig, R, seed=9)
    rb.steps = None
    model.train()
    base = model.get_outputs(rb)["rgb"].detach()
    f.add_table(3)
    f.set_stage("block_stage", active_block=3)
    # a zero residual table leaves the output unchanged (nerfacto_field.py:477-489)
    out = model.get_outputs(rb)["rgb"]
    assert torch.equal(out.detach(), base)
    out.sum().backward()
    res = f.base_encoding_3.hash_3d.feat_pool_
    assert res.grad is not None and float(res.grad.abs().max()) > 0
    assert f.base_encoding_init.hash_3d.feat_pool_.grad is None
    assert all(l.weight.grad is None for l in f.base_network.linears() + f.mlp_head.linears())
    assert f.embedding_appearance.embedding.weight.grad is None
    # a non-zero residual changes the result
    res.data.uniform_(-0.3, 0.3)
    assert not torch.equal(model.get_outputs(rb)["rgb"].detach(), base)
    # table swap to disk and back (save_table / del_table / load_table, nerfacto_field.py:310-403)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        from pathlib import Path
        f.encodings_ckpt_dir = Path(d)
        keep = res.detach().clone()
        f.save_table(3)
        f.del_table(3)
        assert not hasattr(f, "base_encoding_3")
        f.load_table(3)
        assert torch.equal(f.base_encoding_3.hash_3d.feat_pool_.detach(), keep)


def test_engine_block_stage_equals_operator_api_block_stage():
    """Focal stage of the fused engine (frozen global table / MLP / embedding, private residual table, no exchange)
    against the operator-API field in block_stage on the same rays."""
    rig = load_rig("rig8")
    eng, model = build(rig)
    R = 256
    rb, o, d, cam = bundle(rig, R, seed=21)
    rb.steps = None
    target = torch.rand(R, 3, generator=torch.Generator().manual_seed(2)).cuda()
    eng.start_block_stage(lr=5e-3)
    f = model.field
    f.add_table(0)
    f.set_stage("block_stage", active_block=0)
    f.base_encoding_0.load_states(eng.res.States(), 0)          # same primes (and zeros)
    model.train()
    model.persampler.sampler.UpdateMode(1)
    table0, mlp0 = eng.enc.feat_pool_.detach().clone(), eng.mlp.clone()
    out = eng.train_step(o, d, target, cam, optimizer_step=False)
    res = model.get_outputs(rb)
    assert torch.allclose(res["rgb"], out.rgb, rtol=1e-4, atol=1e-5)
    diff = res["rgb"] - target
    (torch.sqrt(diff * diff + 1e-12).sum() / R).backward()
    g = f.base_encoding_0.hash_3d.feat_pool_.grad
    ref = eng.opt_res.unscaled_grad().view_as(g)
    assert float(ref.abs().max()) > 0
    assert float((g - ref).abs().max()) <= 2e-3 * float(ref.abs().max())
    assert not eng.opt_table.grad.any() and not eng.opt_mlp.grad.any()      # frozen: nothing accumulated
    # a few optimizer steps: only the residual table moves, the loss goes down
    losses = []
    for it in range(12):
        losses.append(float(eng.train_step(o, d, target, cam).loss))
    assert losses[-1] < losses[0]
    assert torch.equal(eng.enc.feat_pool_.detach(), table0) and torch.equal(eng.mlp, mlp0)
    assert float(eng.res.feat_pool_.abs().max()) > 0
    assert torch.equal(eng.res._shadow, eng.res.feat_pool_.detach().half())
    eng.end_block_stage()
    assert eng.stage == "init_stage" and eng.res is None


def test_state_dict_keys_match_reference_names():
    rig = load_rig("rig8")
    _, model = build(rig)
    keys = set(model.state_dict().keys())
    for k in ("field.base_network.layers.0.weight", "field.base_network.layers.1.bias", "field.mlp_head.layers.2.weight",
              "field.embedding_appearance.embedding.weight", "field.base_encoding_init.feat_pool",
              "field.base_encoding_init.prime_pool", "field.base_encoding_init.bias_pool",
              "field.base_encoding_init.n_volumes", "persampler.tree_nodes_gpu", "persampler.pers_trans_gpu",
              "persampler.tree_visit_cnt", "persampler.milestones_ts"):
        assert k in keys, (k, sorted(keys))


def test_get_weights_f2nerf_matches_torch_formula_with_autograd():
    import gfnerf_b200 as gf
    R, S = 37, 1024
    g = torch.Generator().manual_seed(0)
    counts = torch.randint(0, S + 1, (R,), generator=g)
    mask = (torch.arange(S)[None] < counts[:, None]).float()
    delta = (torch.rand(R, S, generator=g) * 0.02 * mask).cuda().unsqueeze(-1)
    dens = (torch.rand(R, S, generator=g) * 30).cuda().unsqueeze(-1).requires_grad_(True)
    z = torch.zeros(R, S, 3).cuda()
    rs = gf.RaySamples(frustums=gf.Frustums(z, z, delta, delta), deltas=delta)
    w, a, T = rs.get_weights_f2nerf(dens)
    gw = torch.randn(R, S, 1, generator=g).cuda()
    (w * gw).sum().backward()
    got = dens.grad.clone()
    dens.grad = None
    dd = delta * dens                                          # nerfstudio/cameras/rays.py:188-200
    al = 1 - torch.exp(-dd)
    tr = torch.cumsum(dd[..., :-1, :], dim=-2)
    tr = torch.exp(-torch.cat([torch.zeros(R, 1, 1).cuda(), tr], dim=-2))
    w_ref = torch.nan_to_num(al * tr)
    (w_ref * gw).sum().backward()
    assert torch.allclose(w, w_ref, rtol=1e-5, atol=1e-7) and torch.allclose(a, al, rtol=1e-5, atol=1e-7)
    assert torch.allclose(T, tr, rtol=1e-5, atol=1e-7)
    assert float((got - dens.grad).abs().max()) <= 1e-5 * float(dens.grad.abs().max())


def test_eval_mode_block_routing():
    """Render-time routing of the reference (gfnerf/perssampler.py:138-165, 244-260, 367-376, 429-432): without
    `steps` the chunk's block is the label of the training camera nearest to its first ray origin (or its index
    bucket when no clustering exists), and without `rel_camera_indices` every sample takes that camera's embedding."""
    import gfnerf_b200 as gf
    rig = load_rig("rig8")
    eng, model = build(rig)
    ps = model.persampler
    n_cams = rig["c2w"].shape[0]
    rb, o, d, cam = bundle(rig, 64, seed=9)
    k = 37
    origin = torch.from_numpy(rig["c2w"][k, :, 3]).cuda()
    rb_eval = gf.RayBundle(origins=origin[None].repeat(64, 1) + 1e-3, directions=rb.directions,
                           lookat_directions=rb.directions, pixel_area=rb.pixel_area, camera_indices=rb.camera_indices)
    rs = ps.generate_ray_samples(rb_eval)
    assert rs.cur_step == -1
    assert rs.cur_split_dataset_idx == min(k // (n_cams // ps.n_split_dataset), ps.n_split_dataset - 1)
    assert rs.rel_camera_indices.shape == (64, 1024, 1) and int(rs.rel_camera_indices.min()) == k == int(rs.rel_camera_indices.max())
    labels = torch.arange(n_cams).view(-1, 1) % 8
    ps.cameras_labels = labels.cuda()
    rs = ps.generate_ray_samples(rb_eval)
    assert rs.cur_split_dataset_idx == int(labels[k]) and ps.get_nearest_split_dataset(origin) == (int(labels[k]), k)
    # training-time schedule (:361-364): the block follows the step counter after the init stage
    rb_eval.steps = torch.full((64, 1), ps.steps_perssampler_init + 3 * ps.steps_per_split_dataset + 5).cuda()
    rb_eval.rel_camera_indices = rb.rel_camera_indices
    assert ps.generate_ray_samples(rb_eval).cur_split_dataset_idx == 3
    # the routed bundle renders
    model.eval()
    rb_eval.steps, rb_eval.rel_camera_indices = None, None
    out = model.get_outputs(rb_eval)
    assert torch.isfinite(out["rgb"]).all()


def test_engine_from_model_and_back():
    """`GFNeRFEngine.from_model / load_model / store_model`: the bridge between the operator-API modules (whose state
    dict is the reference's checkpoint layout) and the fused engine.  Same parameters -> same render; a training step
    of the engine goes back into the modules bit for bit."""
    from gfnerf_b200.engine import GFNeRFEngine
    rig = load_rig("rig8")
    eng, model = build(rig, seed=3)
    f = model.field
    f.base_encoding_init.hash_3d.bias_pool_.uniform_(0, 0.1)             # a checkpoint of the reference may carry one
    eng2 = GFNeRFEngine.from_model(model, seed=99)                     # another seed: everything must come from the model
    assert torch.equal(eng2.mlp, eng.mlp) and torch.equal(eng2.emb, eng.emb)
    assert torch.equal(eng2.enc.feat_pool_, eng.enc.feat_pool_) and torch.equal(eng2.enc.prim_pool_, eng.enc.prim_pool_)
    assert torch.equal(eng2.enc.bias_pool_, f.base_encoding_init.hash_3d.bias_pool_)
    R = 256
    rb, o, d, cam = bundle(rig, R, seed=11)
    model.eval()
    res = model.get_outputs(rb)
    out = eng2.render(o, d, cam)
    assert torch.allclose(res["rgb"], out.rgb, rtol=1e-4, atol=1e-5)
    assert torch.allclose(res["depth"][:, 0], out.depth, rtol=1e-4, atol=1e-5)
    # one training step in the engine, then back into the modules
    target = torch.rand(R, 3, generator=torch.Generator().manual_seed(2)).cuda()
    before = eng2.mlp.clone()
    eng2.train_step(o, d, target, cam, update_octree=False)
    eng2.flush()
    assert not torch.equal(eng2.mlp, before) and eng2.opt_mlp.t == 1
    eng2.store_model(model)
    got = torch.cat([f.base_network.flat_params(), f.mlp_head.flat_params()]).detach()
    assert torch.equal(got, eng2.mlp)
    assert torch.equal(f.base_encoding_init.hash_3d.feat_pool_.detach(), eng2.enc.feat_pool_)
    assert torch.equal(f.embedding_appearance.embedding.weight.detach(), eng2.emb)
    res2 = model.get_outputs(rb)
    out2 = eng2.render(o, d, cam)
    assert torch.allclose(res2["rgb"], out2.rgb, rtol=1e-4, atol=1e-5) and not torch.allclose(res2["rgb"], res["rgb"])
    # load_model again: the optimizer restarts
    eng2.load_model(model)
    assert eng2.opt_mlp.t == 0 and not eng2.opt_mlp.m.any()
    # mismatches are refused, not papered over
    _, other = build(rig, log2T=13)
    with pytest.raises(ValueError):
        eng2.load_model(other)
