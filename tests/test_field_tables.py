"""Host logic of the focal-stage table swap of GFNeRFField (add / del / save / load_table, update_active_blocks;
reference gfnerf/nerfacto_field.py:248-403) with a stand-in encoder, so that it runs without a GPU: file format (the
reference's: the encoder's un-prefixed state dict), which tables are resident after each call, which are trainable."""
import os
from pathlib import Path

import torch
from torch import nn


class FakeEncoder(nn.Module):
    """The methods of gfnerf_b200.Hash3DAnchored the field uses, on CPU tensors."""

    def __init__(self, log2_table_size, n_volumes, generator=None):
        super().__init__()
        self.feat = torch.rand(16 << log2_table_size, 2)
        self.prim = torch.arange(16 * n_volumes * 3, dtype=torch.int32).view(16, n_volumes, 3)
        self.bias = torch.zeros(16 * n_volumes, 3)
        self.n_vol = n_volumes
        self.trainable, self.released, self.hooked = None, False, True

    def zero(self):
        self.feat.zero_()

    def state_dict(self, *a, **k):
        return {"feat_pool": self.feat, "prime_pool": self.prim, "bias_pool": self.bias,
                "n_volumes": torch.full((1,), self.n_vol, dtype=torch.int32)}

    def load_state_dict(self, sd, strict=True, prefix=""):
        self.load_states([sd["feat_pool"], sd["prime_pool"], sd["bias_pool"], sd["n_volumes"]], 0)

    def load_states(self, states, idx):
        self.feat, self.prim, self.bias, self.n_vol = states[idx].clone(), states[idx + 1].clone(), states[idx + 2].clone(), int(states[idx + 3].item())
        return idx + 4

    def set_require_grad(self, flag):
        self.trainable = flag

    def unregister_hooks(self):
        self.hooked = False

    def release_resources(self):
        self.released = True


def make_field(tmp_path, monkeypatch, n_blocks=4):
    from gfnerf_b200 import field as fmod
    monkeypatch.setattr(fmod, "Hash3DAnchored", FakeEncoder)
    f = fmod.GFNeRFField.__new__(fmod.GFNeRFField)
    nn.Module.__init__(f)
    f.log2_table_size, f.n_volumes, f.n_blocks = 4, 3, n_blocks
    f.encodings_ckpt_dir = Path(tmp_path) / "encodings_ckpt"
    f.active_block_idxs, f.active_block_idxs_test = [], []
    return f


def test_table_swap_roundtrip_in_the_reference_file_format(tmp_path, monkeypatch):
    f = make_field(tmp_path, monkeypatch)
    f.add_table(2)
    assert not f.base_encoding_2.feat.any()                      # zero-initialised residual (nerfacto_field.py:345)
    f.base_encoding_2.feat.uniform_(-1, 1)
    keep = f.base_encoding_2.feat.clone()
    path = f.save_table(2)
    assert path.endswith("encodings_ckpt/base_encoding_2.ckpt")
    on_disk = torch.load(path)
    assert sorted(on_disk) == ["bias_pool", "feat_pool", "n_volumes", "prime_pool"]     # encoding_field.state_dict()
    enc = f.base_encoding_2
    f.del_table(2)
    assert not hasattr(f, "base_encoding_2") and enc.released and not enc.hooked
    f.load_table(2)
    assert torch.equal(f.base_encoding_2.feat, keep)
    # a file the REFERENCE wrote (same dict) loads; so does the list this repository wrote before
    torch.save({k: v for k, v in on_disk.items()}, str(f.encodings_ckpt_dir / "base_encoding_1.ckpt"))
    f.load_table(1, strict=True)
    assert torch.equal(f.base_encoding_1.feat, keep)
    torch.save([on_disk[k] for k in ("feat_pool", "prime_pool", "bias_pool", "n_volumes")],
               str(f.encodings_ckpt_dir / "base_encoding_3.ckpt"))
    f.load_table(3)
    assert torch.equal(f.base_encoding_3.feat, keep)
    # a missing file: silently nothing unless strict (nerfacto_field.py:393-403)
    f.load_table(0)
    assert not hasattr(f, "base_encoding_0")
    try:
        f.load_table(0, strict=True)
        raise AssertionError("strict load of a missing table must fail")
    except FileNotFoundError:
        pass


def test_update_active_blocks_keeps_only_the_active_tables_resident(tmp_path, monkeypatch):
    f = make_field(tmp_path, monkeypatch)
    f.train()
    f.update_active_blocks(1)                                    # block 1 becomes the training block
    assert f.active_block_idxs == [1] and hasattr(f, "base_encoding_1") and f.base_encoding_1.trainable is True
    f.base_encoding_1.feat.fill_(0.5)
    f.update_active_blocks(2)                                    # switch: block 1 is swapped out to disk
    assert f.active_block_idxs == [2] and not hasattr(f, "base_encoding_1") and hasattr(f, "base_encoding_2")
    assert os.path.exists(f.encodings_ckpt_dir / "base_encoding_1.ckpt")
    f.eval()
    f.update_active_blocks(1)                                    # eval on block 1: loaded back (must exist), frozen
    assert f.active_block_idxs_test == [1] and f.active_block_idxs == [2]
    assert float(f.base_encoding_1.feat.min()) == 0.5 and f.base_encoding_1.trainable is False
    assert hasattr(f, "base_encoding_2")                         # the training block stays resident
    f.update_active_blocks(-1)                                   # no eval block any more
    assert f.active_block_idxs_test == [] and not hasattr(f, "base_encoding_1") and hasattr(f, "base_encoding_2")


# ---- MLPNetwork: parameter names of the reference's own class (gfnerf/mlp.py:35-43) ---------------------------------

REF_MLP_KEYS = {  # what `MLPNetwork(...).state_dict()` of the reference yields for the two stacks of the field
    "base": ["layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias"],
    "head": ["layers.0.weight", "layers.0.bias", "layers.1.weight", "layers.1.bias", "layers.2.weight",
             "layers.2.bias"],
}


def _stacks(cls, hidden=64):
    cfg = lambda act, n: {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": act,
                          "n_neurons": hidden, "n_hidden_layers": n}
    return cls(32, 16, cfg("None", 1)), cls(63, 3, cfg("Sigmoid", 2))


def test_mlpnetwork_state_dict_keys_are_the_references():
    from gfnerf_b200.mlp import MLPNetwork
    base, head = _stacks(MLPNetwork)
    assert list(base.state_dict().keys()) == REF_MLP_KEYS["base"]
    assert list(head.state_dict().keys()) == REF_MLP_KEYS["head"]


def test_mlpnetwork_loads_a_state_dict_of_the_references_class():
    """strict load of a state dict produced by the reference's own MLPNetwork (where the reference tree exists)"""
    import importlib.util
    import pytest
    path = "/root/reference/gfnerf/mlp.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    spec = importlib.util.spec_from_file_location("ref_gfnerf_mlp", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from gfnerf_b200.mlp import MLPNetwork
    for H in (64, 128):
        for mine, theirs in zip(_stacks(MLPNetwork, H), _stacks(ref.MLPNetwork, H)):
            assert list(theirs.state_dict().keys()) == list(mine.state_dict().keys())
            mine.load_state_dict(theirs.state_dict(), strict=True)
            theirs.load_state_dict(mine.state_dict(), strict=True)
            ref_flat = torch.cat([torch.cat([l.weight.reshape(-1), l.bias.reshape(-1)]) for l in theirs.layers])
            assert torch.equal(mine.flat_params(), ref_flat)
