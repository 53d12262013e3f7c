"""Fused-kernel rate for a few variants (device-resident sums, CUDA events on the handle's stream)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
only = sys.argv[2] if len(sys.argv) > 2 else ""            # substring filter on the case names
h = _lib.Handle(0)
g = SVJParams.gbm(0.3, r=0.065)
bumps = _lib.Bumps(0.01, g.v0 + 0.01, g.v0 - 0.01, g.r + 1e-4, g.r - 1e-4)
out = h.malloc(17 * 8 * 256)
gk = SVJParams(kappa=3.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.0, r=0.065, q=0.0)
cases = [("gbm fp32 greeks", g, _lib.GREEKS, bumps, [2500.0], 2500.0, n),
         ("kappa=3 fp32 greeks (detvar)", gk, _lib.GREEKS, bumps, [2500.0], 2500.0, n),
         ("kappa=3 fp32 price (gbm)", gk, 0, None, [2500.0], 2500.0, n),
         ("gbm fp32 price", g, 0, None, [2500.0], 2500.0, n),
         ("gbm fp32 anti", g, _lib.ANTITHETIC, None, [2500.0], 2500.0, n),
         ("gbm fp64 price", g, _lib.FP64, None, [2500.0], 2500.0, n),
         ("gbm fp64 greeks", g, _lib.FP64 | _lib.GREEKS, bumps, [2500.0], 2500.0, n),
         ("gbm fp64 anti", g, _lib.FP64 | _lib.ANTITHETIC, None, [2500.0], 2500.0, n),
         ("gbm fp32 64 strikes", g, _lib.ANTITHETIC, None, list(np.linspace(0.7, 1.3, 64) * 2500.0), 2500.0, n),
         ("heston fp32 anti", SVJParams(lambda_j=0.0), _lib.ANTITHETIC, None, [22500.0], 22500.0, n // 4),
         ("svj fp32 anti", SVJParams(), _lib.ANTITHETIC, None, [22500.0], 22500.0, n // 4),
         ("svj fp32 greeks", SVJParams(), _lib.GREEKS, _lib.Bumps(0.01, 0.05, 0.03, 0.0651, 0.0649), [22500.0], 22500.0, n // 4)]
print("library:", _lib.LIB_PATH, flush=True)
for name, p, fl, b, ks, s0, nn in cases:
    if only not in name:
        continue
    best = 1e9
    for r in range(4):
        h.timer_begin()
        h.price_european(p, s0, 1.0, 250, nn, 42 + r, ks, True, fl, b, out_dev=out)
        ms = h.timer_end()
        if r:
            best = min(best, ms)
    print(f"{name:28s} {nn:9d} paths x 250: {best:8.3f} ms  {nn * 250 / best / 1e9:8.2f} Gpath-steps/ms -> {nn * 250 / best / 1e6:9.1f} G/s", flush=True)
h.close()
