"""Small-call latency through the public API (what calibration.py / the web handlers see)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import GreeksEngine, MonteCarloEngine, SVJParams  # noqa: E402

p = SVJParams()
for n, steps in ((10_000, 50), (100_000, 50), (50_000, 252), (500_000, 252)):
    e = MonteCarloEngine(p, n, steps, 42, use_sobol=False)
    e.price(22500.0, 22500.0, 1.0)
    t0 = time.perf_counter()
    for i in range(200):
        e.price(22500.0, 22000.0 + i, 1.0)
    dt = (time.perf_counter() - t0) / 200
    print(f"price()  {n:7d} paths x {steps:3d} steps (SVJ, antithetic+CV): {dt * 1e6:8.1f} us/call")
g = GreeksEngine(p, 50_000, 252, 42)
g.delta(22500.0, 22500.0, 0.25)
t0 = time.perf_counter()
for i in range(100):
    g.seed = i
    g.delta(22500.0, 22500.0, 0.25); g.vega(22500.0, 22500.0, 0.25); g.gamma(22500.0, 22500.0, 0.25)
print(f"delta+vega+gamma 50k x 63 (one fused launch): {(time.perf_counter() - t0) / 100 * 1e6:8.1f} us")
e = MonteCarloEngine(p, 50_000, 252, 42, use_sobol=False)
t0 = time.perf_counter()
for i in range(50):
    e.price_batch(22500.0, [22500.0 * (0.7 + 0.03 * k) for k in range(21)], 0.25)
print(f"price_batch 21 strikes 50k x 63: {(time.perf_counter() - t0) / 50 * 1e6:8.1f} us")
t0 = time.perf_counter()
for i in range(50):
    e.get_sample_paths(22500.0, 0.25, 50)
print(f"get_sample_paths 50 x 64: {(time.perf_counter() - t0) / 50 * 1e6:8.1f} us")
