"""Quasi-Monte Carlo front end (rng="sobol") vs plain Monte Carlo (rng="philox"): RMS error against Black-Scholes over
8 seeds for a GBM call (S0 = K = 2500, T = 1, sigma = 0.3, 250 steps), and the time per price() call."""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams, _lib, bs_price  # noqa: E402

h = _lib.Handle(0)
p = SVJParams.gbm(0.3, r=0.065)
bs = bs_price(2500.0, 2500.0, 1.0, 0.065, 0.0, 0.3, True)
print(f"Black-Scholes {bs:.6f}")
for n in (1024, 4096, 16384, 65536, 262144, 1048576):
    row = []
    for rng in ("sobol", "philox"):
        errs, t = [], 0.0
        for seed in range(8):
            e = MonteCarloEngine(p, n, 250, seed, use_antithetic=False, use_control_variate=False, rng=rng, handle=h)
            e.price(2500.0, 2500.0, 1.0)
            t0 = time.perf_counter()
            r = e.price(2500.0, 2500.0, 1.0)
            t += time.perf_counter() - t0
            errs.append(r["price"] - bs)
        row.append(f"{rng}: rms error {math.sqrt(np.mean(np.square(errs))):9.5f}  {t / 8 * 1e3:7.2f} ms/call")
    print(f"n = {n:8d}   " + "   ".join(row), flush=True)
h.close()
