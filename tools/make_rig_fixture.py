"""Builds the synthetic aerial-rig octree fixtures with the host builder and stores the state
blobs, so that tests and bench.py do not spend minutes in octree construction.

  python tools/make_rig_fixture.py 8  tests/golden/rig8.npz     # parity tests (64 cameras)
  python tools/make_rig_fixture.py 20 tests/golden/rig20.npz    # bench (400 cameras, BASELINE config 2)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util

spec = importlib.util.spec_from_file_location("persoctree", os.path.join(os.path.dirname(__file__), "..", "gf-nerf_b200", "persoctree.py"))
po = importlib.util.module_from_spec(spec)
spec.loader.exec_module(po)


def main():
    n_side, out = int(sys.argv[1]), sys.argv[2]
    extent = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
    c2w, intri, bounds = po.aerial_rig(n_side=n_side, extent=extent, seed=1)
    t = time.time()
    oc = po.PersOctree(16, 512.0, 1.5, c2w, intri, bounds, seed=0)
    print(f"built in {time.time() - t:.1f}s: {oc.nodes.shape[0]} nodes, {oc.trans.shape[0]} transforms")
    np.savez_compressed(out, c2w=c2w, intri=intri, bounds=bounds, tree_nodes=oc.tree_nodes_blob(),
                        pers_trans=oc.pers_trans_blob(), n_side=n_side, extent=extent)


if __name__ == "__main__":
    main()
