#!/usr/bin/env python
"""Generates tests/golden/ref_marks.npz by RUNNING THE REFERENCE's own MarkInvisibleNodesKernel / CheckVisible and
SetBlockIdxsNearestKernel (gfnerf/bindings/PtsSampler/PersSampler_cuda.cu:680-766) -- extracted where they lie under
/root/reference and compiled for the host by `make -C oracle ref` -- on the rig fixtures' octrees.  The fixture holds
the inputs (camera subsets, block centres) and the reference's outputs (trans_idx / block_idx per node), so the pin
travels to machines without /root/reference: tests/test_ref_kernels.py (oracle, CPU) and
tests/test_octree_device_gpu.py (gf_octree_mark_invisible / gf_octree_set_block_idxs, GPU).  Both contraction
flavours of the host build give the same integers on these inputs (asserted below).

  make -C oracle ref && python tests/golden/make_golden_marks.py      # build container only
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_host as rh  # noqa: E402
from tests.helpers import load_rig  # noqa: E402


def w2c_of(c2w):
    m = np.tile(np.eye(4, dtype=np.float32)[None], (c2w.shape[0], 1, 1))
    m[:, :3, :] = c2w
    return np.ascontiguousarray(np.linalg.inv(m)[:, :3, :].astype(np.float32))


def main():
    fx = {}
    for name, cams in (("rig8", (0, 1, 2, 3, 4, 5)), ("rig20", tuple(range(0, 400, 37)))):
        rig = load_rig(name)
        sel = np.array(cams)
        w2c = w2c_of(rig["c2w"][sel])
        intri, bounds = np.ascontiguousarray(rig["intri"][sel]), np.ascontiguousarray(rig["bounds"][sel])
        centers = np.random.RandomState(len(cams)).uniform(-4, 4, size=(5, 3)).astype(np.float32)
        centers = np.concatenate([centers, centers[1:3]], 0)           # exact ties: blocks 5, 6 duplicate 1, 2
        out = {}
        for flavour in ("off", "fma"):
            nodes = rig["tree_nodes"].copy()
            rh.mark_invisible_nodes(nodes, intri, w2c, bounds, flavour=flavour)
            rh.set_block_idxs(nodes, centers, flavour=flavour)
            out[flavour] = nodes
        assert np.array_equal(out["off"], out["fma"]), "the two host flavours disagree: pick other inputs"
        blob = out["fma"].view(np.int64).reshape(-1, 16)
        fx.update({f"{name}_w2c": w2c, f"{name}_intri": intri, f"{name}_bounds": bounds, f"{name}_centers": centers,
                   f"{name}_trans_idx": blob[:, 12].copy(), f"{name}_block_idx": blob[:, 13].copy()})
        before = rig["tree_nodes"].view(np.int64).reshape(-1, 16)[:, 12]
        print(name, "nodes", blob.shape[0], "cameras", len(cams), "lost their transform:",
              int(((blob[:, 12] == -1) & (before != -1)).sum()), "of", int((before != -1).sum()))
    path = os.path.join(HERE, "ref_marks.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB")


if __name__ == "__main__":
    main()
