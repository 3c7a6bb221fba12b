"""The random stream of the native NUTS (csrc/sfh_nuts.h) as a numpy-Generator-shaped object (test infrastructure): Philox4x32-10
keyed by seed, counter = (chain << 40 | draw index), stream 32; normals by Box-Muller from two uniforms each, evaluated with the
scalar libm functions the C++ side uses.  Handing it to `sfh_b200.solvers.nuts_chain` (the Python engine, same algorithm, same
order of draws) yields the executable restatement of a native chain."""
import math

import numpy as np

from ensemble_ref import philox_u01

DRAW_NUTS = 32


class PhiloxRng:
    def __init__(self, seed, chain):
        self.seed, self.base, self.n = int(seed), int(chain) << 40, 0

    def random(self):
        u = float(philox_u01(np.array([self.base | self.n], dtype=np.uint64), self.seed, DRAW_NUTS)[0])
        self.n += 1
        return u

    def standard_normal(self, d):
        out = np.empty(d)
        for i in range(d):
            u1, u2 = self.random(), self.random()
            out[i] = math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(6.283185307179586 * u2)
        return out
