"""(cell x path) grid launch vs a loop of single-problem launches: the hedging-backtest premiums (1000 scenarios x 50k
paths x 63 steps, engine/risk.py:264-273) and the 13 scenarios of a stress report (200k paths)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

h = _lib.Handle(0)
for name, p in (("gbm", SVJParams.gbm(0.3, r=0.065)), ("svj", SVJParams())):
    for n_cells, n_paths, steps in ((1000, 50_000, 63), (13, 200_000, 63), (64, 1_000_000, 250)):
        cells = _lib.make_cells(p, 22500.0, steps / 252, steps, n_paths, 42 + np.arange(n_cells))
        ks = np.full(n_cells, 22500.0)
        out = h.malloc(n_cells * 17 * 8)
        best_b = best_l = best_w = 1e9
        for r in range(4):
            t0 = time.perf_counter()
            h.timer_begin()
            h.price_cells(cells, ks, _lib.ANTITHETIC, out_dev=out)
            ms = h.timer_end()
            w = (time.perf_counter() - t0) * 1e3
            if r:
                best_b, best_w = min(best_b, ms), min(best_w, w)
        for r in range(3):
            h.timer_begin()
            for i in range(n_cells):
                h.price_european(p, 22500.0, steps / 252, steps, n_paths, 42 + i, [22500.0], True, _lib.ANTITHETIC,
                                 out_dev=out + i * 17 * 8)
            ms = h.timer_end()
            if r:
                best_l = min(best_l, ms)
        h.free(out)
        work = n_cells * n_paths * steps
        print(f"{name} {n_cells:5d} cells x {n_paths:8d} paths x {steps:3d}: one launch {best_b:8.3f} ms "
              f"({work / best_b / 1e9:6.3f}e12 path-steps/s, wall {best_w:7.2f} ms)   "
              f"loop of launches {best_l:8.3f} ms", flush=True)

# BASELINE cfg3, benchmark reading (64 strikes x 16 expiries, 1M paths per cell, own draws per cell) as ONE launch
p = SVJParams.gbm(0.3, r=0.065)
Ts = np.repeat(np.arange(1, 17) / 8.0, 64)
steps = np.maximum((250 * Ts).astype(int), 10)
order = np.argsort(-steps, kind="stable")                     # longest cells first
cells = _lib.make_cells(p, 2500.0, Ts, steps, 1_000_000, 42, np.arange(1024) * 1_000_000)[order]
ks = np.tile(np.linspace(0.7, 1.3, 64) * 2500.0, 16)[order]
out = h.malloc(1024 * 17 * 8)
best = 1e9
for r in range(3):
    h.timer_begin()
    h.price_cells(cells, ks, _lib.ANTITHETIC, out_dev=out)
    best = min(best, h.timer_end())
h.free(out)
print(f"cfg3 independent cells as one launch: {best:.2f} ms = {float(steps.sum()) * 1e6 / best / 1e9:.3f}e12 path-steps/s", flush=True)

# the two callers end to end (wall clock, Python included)
from monte_carlo_option_simulator_b200.risk import HedgingBacktest, StressTestEngine  # noqa: E402
for name, p in (("gbm", SVJParams.gbm(0.3, r=0.065)), ("svj", SVJParams())):
    st = StressTestEngine(p, num_paths=200_000, seed=42, handle=h)
    bt = HedgingBacktest(p, seed=42, handle=h)
    st.full_stress_report(22500.0, 22500.0, 0.25)
    bt.run_backtest(22500.0, 22500.0, 0.25, num_scenarios=10, num_mc_paths=1000)
    t0 = time.perf_counter()
    for _ in range(5):
        st.full_stress_report(22500.0, 22500.0, 0.25)
    t_st = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    res = bt.run_backtest(22500.0, 22500.0, 0.25)
    t_bt = time.perf_counter() - t0
    print(f"{name}: full_stress_report (11 cells x 200k paths x 63) {t_st * 1e3:.2f} ms;  run_backtest (1000 scenarios x 50k paths "
          f"x 63 + walk) {t_bt * 1e3:.1f} ms  mean_pnl {res['mean_pnl']:.2f} std {res['std_pnl']:.2f}", flush=True)
h.close()
