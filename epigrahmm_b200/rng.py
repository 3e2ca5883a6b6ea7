"""R's default RNG (Mersenne-Twister + set.seed scrambling) on the host.

The reference's rejection control draws from R's global stream through ``RNGScope``
(src/RcppExports.cpp:19; src/rbinomVectorized.cpp:17), so the stream state is an input and an
output of the hot path.  This module holds that state for the host-side mirror: ``set_seed``
reproduces ``set.seed`` (LCG scrambling 69069*s+1: 50 warm-up steps, then 625 words, mti = 624) and
the state array has the layout of ``.Random.seed[2:626]`` (mti, then the 624 words), which is what
``ehmm_rng_set_state`` / ``ehmm_rng_get_state`` exchange with the device generator.
numpy's MT19937 bit generator supplies the recurrence for host-side draws.
"""
from __future__ import annotations

import numpy as np

_I2_32M1 = 2.328306437080797e-10


class RState:
    def __init__(self, seed: int | None = None):
        self.state = np.zeros(625, dtype=np.uint32)
        self.state[0] = 624
        if seed is not None:
            self.set_seed(seed)

    def set_seed(self, seed: int) -> "RState":
        s = np.uint32(seed & 0xFFFFFFFF)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(69069) * s + np.uint32(1)
            for j in range(625):
                s = np.uint32(69069) * s + np.uint32(1)
                self.state[j] = s
        self.state[0] = 624
        return self

    def runif(self, n: int) -> np.ndarray:
        """n draws of unif_rand(): MT_genrand() * 2^-32 with R's fixup into (0,1)."""
        bg = np.random.MT19937()
        bg.state = {"bit_generator": "MT19937", "state": {"key": self.state[1:].copy(), "pos": int(self.state[0])}}
        raw = bg.random_raw(int(n)).astype(np.float64)
        st = bg.state["state"]
        self.state[1:] = st["key"]
        self.state[0] = st["pos"]
        u = raw * 2.3283064365386963e-10
        u[u <= 0.0] = 0.5 * _I2_32M1
        u[(1.0 - u) <= 0.0] = 1.0 - 0.5 * _I2_32M1
        return u


GLOBAL = RState(seed=0)


def set_seed(seed: int) -> None:
    """set.seed(seed) for the module-level stream used by epigrahmm_b200.api"""
    GLOBAL.set_seed(seed)
