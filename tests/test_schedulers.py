"""gfnerf_b200.schedulers against multipliers produced by the reference's GFNerfExponentialDecayScheduler
(tests/golden/make_golden.py runs nerfstudio/engine/schedulers.py:138-184 through torch's LambdaLR).  CPU only."""
import ast
import os

import numpy as np

from gfnerf_b200.schedulers import GFNerfExponentialDecaySchedulerConfig, gfnerf_exponential_decay

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_scheduler.npz")


def test_multipliers_match_reference_scheduler():
    g = np.load(GOLD)
    for name in ("init", "block"):
        cfg = GFNerfExponentialDecaySchedulerConfig(**ast.literal_eval(str(g[name + "_cfg"])))
        got = np.array([gfnerf_exponential_decay(int(s), 1e-2, cfg) for s in g["steps"]])
        np.testing.assert_allclose(got, g[name + "_mult"], rtol=1e-12, atol=0)
    # decays from 1 to lr_final / lr_init over the init stage, restarts at every split dataset of the block stage
    cfg = GFNerfExponentialDecaySchedulerConfig(lr_final=1e-4, max_steps=1000, steps_perssampler_init=1000,
                                                steps_per_split_dataset=500, n_split_dataset=2)
    near = lambda a, b: abs(a - b) < 1e-12
    assert near(gfnerf_exponential_decay(0, 1e-2, cfg), 1.0)
    assert near(gfnerf_exponential_decay(999, 1e-2, cfg), 10 ** (-2 * 0.999))
    assert near(gfnerf_exponential_decay(1000, 1e-2, cfg), 1.0) and near(gfnerf_exponential_decay(1500, 1e-2, cfg), 1.0)
