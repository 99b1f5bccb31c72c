"""Learning-rate policy (reference src/Optim/lr_utils.py:9-32): the factor LambdaLR multiplies Adam's lr = 1.0 with."""
from __future__ import annotations

import math
from dataclasses import dataclass


def _unit_clamp(x: float) -> float:
    return 0.0 if x < 0.0 else (1.0 if x > 1.0 else x)


@dataclass(frozen=True)
class LRDecayPolicy:
    """lr(i) = warm_up(i) * exp(lerp(log lr_init, log lr_final, i / max_steps)); the warm-up is a quarter sine from
    lr_delay_mult to 1 over the first lr_delay_steps iterations (off when lr_delay_steps == 0)."""
    lr_init: float = 1.0
    lr_final: float = 1.0
    lr_delay_steps: int = 0
    lr_delay_mult: float = 1.0
    max_steps: int = 1_000_000

    def _warm_up(self, iteration: int) -> float:
        if self.lr_delay_steps <= 0 or iteration >= self.lr_delay_steps:
            return 1.0
        quarter_sine = math.sin(0.5 * math.pi * _unit_clamp(iteration / self.lr_delay_steps))
        return self.lr_delay_mult + (1.0 - self.lr_delay_mult) * quarter_sine

    def __call__(self, iteration: int) -> float:
        disabled = self.lr_init == 0.0 and self.lr_final == 0.0
        if disabled or iteration < 0:
            return 0.0
        progress = _unit_clamp(iteration / self.max_steps)
        log_lr = (1.0 - progress) * math.log(self.lr_init) + progress * math.log(self.lr_final)
        return float(self._warm_up(iteration) * math.exp(log_lr))
