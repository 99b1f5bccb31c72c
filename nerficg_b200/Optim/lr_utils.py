"""Learning-rate policy (reference src/Optim/lr_utils.py:9-32)."""
from __future__ import annotations

import math
from dataclasses import dataclass


@dataclass(frozen=True)
class LRDecayPolicy:
    """Log-linear interpolation lr_init -> lr_final over max_steps with an optional sine warm-up."""
    lr_init: float = 1.0
    lr_final: float = 1.0
    lr_delay_steps: int = 0
    lr_delay_mult: float = 1.0
    max_steps: int = 1_000_000

    def __call__(self, iteration: int) -> float:
        if iteration < 0 or (self.lr_init == 0.0 and self.lr_final == 0.0):
            return 0.0
        delay = 1.0
        if self.lr_delay_steps > 0 and iteration < self.lr_delay_steps:
            ramp = min(max(iteration / self.lr_delay_steps, 0.0), 1.0)
            delay = self.lr_delay_mult + (1 - self.lr_delay_mult) * math.sin(0.5 * math.pi * ramp)
        t = min(max(iteration / self.max_steps, 0.0), 1.0)
        return float(delay * math.exp(math.log(self.lr_init) * (1 - t) + math.log(self.lr_final) * t))
