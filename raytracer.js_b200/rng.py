"""RNG classes (src/math/rng/rng.ts, prng.ts, fp-lcg.ts)."""
from __future__ import annotations

import math

PRNG_MUL1 = 3532205053565347.0 / 3768278866164713.0
PRNG_TERM1 = 3773467585272041.0 / 4435662911655887.0
PRNG_MUL2 = 3632519696538149.0 / 4496133748415501.0
PRNG_TERM2 = 3396159042346757.0 / 4429161683464229.0
PRNG_MUL3 = 4056279137291581.0 / 4272384783187219.0
PRNG_TERM3 = 3685311960670787.0 / 3909517015383373.0


class RNG:
    def next(self) -> float:
        raise NotImplementedError


class PRNG(RNG):
    def seed(self, seed: float) -> None:
        raise NotImplementedError


class FpLcg(PRNG):
    """Floating-point tri-state LCG with state mixing (src/math/rng/fp-lcg.ts:49-82)."""

    def __init__(self, seed: float):
        self.seed_value = float(seed)
        self.seed(seed)

    def seed(self, seed: float) -> None:
        self.seed_value = float(seed)
        self.state1 = float(seed)
        self.state2 = seed * PRNG_MUL3
        self.state3 = seed * PRNG_MUL2

    def next(self) -> float:
        s1 = math.fmod(self.state1 * PRNG_MUL1 + PRNG_TERM1, 1.0)
        s2 = math.fmod(self.state2 * PRNG_MUL2 + PRNG_TERM2, 1.0)
        s3 = math.fmod(self.state3 * PRNG_MUL3 + PRNG_TERM3, 1.0)
        self.state1 = s2 + s3
        self.state2 = s3
        self.state3 = s1 + s2
        return math.fmod(s1 + s2 + s3, 1.0)
