"""Launch time of the fused kernel against strike count and path count (device-resident sums, CUDA events)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

h = _lib.Handle(0)
g = SVJParams.gbm(0.3, r=0.065)
out = h.malloc(17 * 8 * 256)
for ns in (1, 8, 21, 64, 256):
    ks = list(np.linspace(0.7, 1.3, ns) * 2500.0)
    row = []
    for n in (256, 151_552, 1_000_000, 10_000_000):
        best = 1e9
        for r in range(5):
            h.timer_begin()
            h.price_european(g, 2500.0, 1.0, 248, n, 42 + r, ks, True, _lib.ANTITHETIC, None, out_dev=out)
            ms = h.timer_end()
            if r:
                best = min(best, ms)
        row.append(f"{n:>9d}: {best * 1e3:8.1f} us")
    print(f"{ns:4d} strikes  " + "  ".join(row), flush=True)
h.close()
