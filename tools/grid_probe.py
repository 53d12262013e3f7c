import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams
p = SVJParams.gbm(0.3, r=0.065)
ks, Ts = np.linspace(0.7, 1.3, 64) * 2500.0, [j / 8 for j in range(1, 17)]
for k in (1, 2, 4, 8):
    os.environ["B200MC_GRID_STREAMS"] = str(k)
    e = MonteCarloEngine(p, 1_000_000, 250, 42, use_sobol=False, use_antithetic=False, use_control_variate=False, rng="philox")
    e.price_grid(2500.0, ks[:2], Ts[:2], True, independent_cells=True)
    t0 = time.perf_counter(); g = e.price_grid(2500.0, ks, Ts, True, independent_cells=True); dt = time.perf_counter() - t0
    print(k, "streams:", round(dt * 1e3, 1), "ms", "%.3e path-steps/s" % (64e6 * g["num_steps"].sum() / dt), g["prices"][7, 32])
