"""Large-sample bias check of the fused GBM kernel against Black-Scholes (log-Euler with constant variance has no
discretisation bias, so any deviation beyond a few standard errors is generator / arithmetic bias).
    python tools/bias_check.py [n_paths]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams, bs_price  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
p = SVJParams.gbm(0.3, r=0.065)
for K in (2500.0, 2000.0, 3200.0):
    bs = bs_price(2500.0, K, 1.0, p.r, p.q, 0.3, True)
    for prec in ("fp32", "fp64"):
        for seed in (42, 7):
            e = MonteCarloEngine(p, n, 250, seed, use_sobol=False, use_antithetic=False, use_control_variate=False,
                                 rng="philox", precision=prec)
            r = e.price(2500.0, K, 1.0, True)
            z = (r["price"] - bs) / r["std_error"]
            zc = (r["price_cv_spot"] - bs) / r["std_error_cv_spot"]
            print(f"K={K:6.0f} {prec} seed={seed:3d} n={n:.0e}: MC {r['price']:.5f} +- {r['std_error']:.5f}  BS {bs:.5f}  "
                  f"z={z:+.2f} rel={abs(r['price'] - bs) / bs:.1e} | spot-CV {r['price_cv_spot']:.5f} +- {r['std_error_cv_spot']:.5f} z={zc:+.2f}",
                  flush=True)
