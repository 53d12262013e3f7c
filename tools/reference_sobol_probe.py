"""rng="reference", use_sobol=True -- the reference's DEFAULT engine configuration -- with the Sobol front end on the
device (default) and on the host (B200MC_REFERENCE_SOBOL=host, the reference's own NumPy/SciPy code path): same numbers,
time per price() call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams, _lib  # noqa: E402

h = _lib.Handle(0)
for name, p in (("svj defaults", SVJParams()), ("heston", SVJParams(lambda_j=0.0))):
    for n in (50_000, 500_000):
        res = {}
        for where in ("device", "host"):
            if where == "host" and n > 50_000:
                continue
            os.environ["B200MC_REFERENCE_SOBOL"] = where
            e = MonteCarloEngine(p, num_paths=n, seed=42, rng="reference", handle=h)       # default flags: Sobol, antithetic, CV
            e.price(22500.0, 22500.0, 0.25)
            t0 = time.perf_counter()
            r = e.price(22500.0, 22500.0, 0.25)
            res[where] = (time.perf_counter() - t0, r["price"], r["raw_mc_price"])
        line = f"{name:13s} n={n:7d} x 63 steps:  device {res['device'][0] * 1e3:8.1f} ms  price {res['device'][1]:.6f} raw {res['device'][2]:.6f}"
        if "host" in res:
            line += f"   host front end {res['host'][0] * 1e3:8.1f} ms  price {res['host'][1]:.6f} raw {res['host'][2]:.6f}"
        print(line, flush=True)

# the heaviest caller in the reference's default configuration: HedgingBacktest (1000 scenarios, each a fresh engine seeded
# seed + scenario with the Sobol front end; the reference spends ~8 s per scenario)
from monte_carlo_option_simulator_b200.risk import HedgingBacktest, StressTestEngine  # noqa: E402
os.environ["B200MC_REFERENCE_SOBOL"] = "device"
for nsc in (20, 1000):
    t0 = time.perf_counter()
    r = HedgingBacktest(SVJParams(), seed=42, rng="reference", handle=h).run_backtest(22500.0, 22500.0, 0.25, num_scenarios=nsc)
    print(f"HedgingBacktest rng=reference, {nsc} scenarios x 50k paths: {time.perf_counter() - t0:.2f} s  mean_pnl {r['mean_pnl']:.4f}", flush=True)
t0 = time.perf_counter()
rep = StressTestEngine(SVJParams(), num_paths=200_000, rng="reference", handle=h).full_stress_report(22500.0, 22500.0, 0.25)
print(f"StressTestEngine rng=reference, 200k paths: {time.perf_counter() - t0:.3f} s  base {rep['jump_scenario']['base_price']:.6f}", flush=True)
h.close()
