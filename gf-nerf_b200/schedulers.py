"""Learning-rate schedule of gf-nerf as a pure function of the step.

Mirrors `GFNerfExponentialDecayScheduler.get_scheduler` (reference nerfstudio/engine/schedulers.py:138-184): linear or
cosine warm-up, then an exponential decay from lr_init to lr_final over `max_steps`, restarted per split dataset in
the block (focal) stage.  Returns the multiplier torch's LambdaLR would apply to lr_init; the fused engine passes it
to `train_step(lr_scale=...)`.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class GFNerfExponentialDecaySchedulerConfig:
    """nerfstudio/engine/schedulers.py:112-136 (field names and defaults of the reference config)"""
    lr_pre_warmup: float = 1e-8
    lr_final: Optional[float] = None
    warmup_steps: int = 0
    max_steps: int = 100000
    ramp: str = "cosine"
    n_split_dataset: int = 1
    n_dataset_circles: int = 1
    steps_per_split_dataset: int = 1000
    steps_perssampler_init: int = 10000


def gfnerf_exponential_decay(step: int, lr_init: float, cfg: GFNerfExponentialDecaySchedulerConfig) -> float:
    lr_final = lr_init if cfg.lr_final is None else cfg.lr_final
    if step < cfg.warmup_steps:
        if cfg.ramp == "cosine":
            lr = cfg.lr_pre_warmup + (1 - cfg.lr_pre_warmup) * np.sin(0.5 * np.pi * np.clip(step / cfg.warmup_steps, 0, 1))
        else:
            lr = cfg.lr_pre_warmup + (lr_init - cfg.lr_pre_warmup) * step / cfg.warmup_steps
    else:
        init = cfg.steps_perssampler_init > 0 and step < cfg.steps_perssampler_init
        if init:
            relative_step = step
        else:
            span = cfg.steps_per_split_dataset * cfg.n_split_dataset
            idx = ((step - cfg.steps_perssampler_init) // cfg.steps_per_split_dataset) % cfg.n_split_dataset
            circles = (step - cfg.steps_perssampler_init) // span
            relative_step = (step - cfg.steps_perssampler_init - circles * span - idx * cfg.steps_per_split_dataset
                             + circles * cfg.steps_per_split_dataset)
        t = np.clip((relative_step - cfg.warmup_steps) / (cfg.max_steps - cfg.warmup_steps), 0, 1)
        lr = np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
    return float(lr / lr_init)
