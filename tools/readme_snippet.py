"""The quick-start snippet of README.md, runnable (needs a B200)."""
import sys; sys.path.insert(0, ".")
from monte_carlo_option_simulator_b200 import MonteCarloEngine, GreeksEngine, SVJParams
from monte_carlo_option_simulator_b200.risk import StressTestEngine, HedgingBacktest
from monte_carlo_option_simulator_b200.surface import iv_surface_from_engine
p = SVJParams()
print(MonteCarloEngine(p, num_paths=10_000_000).price(22500.0, 22500.0, 0.25)["price"])
print(GreeksEngine(p, num_paths=10_000_000).all_greeks(22500.0, 22500.0, 0.25)["delta"])
print(StressTestEngine(p).full_stress_report(22500.0, 22500.0, 0.25)["jump_scenario"])
print(HedgingBacktest(p).run_backtest(22500.0, 22500.0, 0.25)["mean_pnl"])
print(MonteCarloEngine(SVJParams.gbm(0.3), 1 << 20, rng="sobol").price(2500.0, 2500.0, 1.0)["price"])
print(iv_surface_from_engine(MonteCarloEngine(p, 1_000_000, use_control_variate=False), 22500.0, [21000.0, 22500.0, 24000.0], [0.1, 0.25, 0.5])["iv_call"])
