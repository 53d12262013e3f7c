"""Golden outputs of the reference's StressTestEngine and HedgingBacktest (engine/risk.py:23-111, 238-337), written by
running the reference itself in the dev container:

    python tests/golden/make_risk_callers_golden.py        ->  tests/golden/risk_callers_golden.json

tests/test_oracle.py pins oracle.StressOracle / oracle.HedgingOracle to these; the GPU tests then compare the CUDA path
(rng="reference": the reference's own host draws) with the same numbers.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, "/root/reference")

from engine.models import SVJParams                      # noqa: E402  (the reference itself)
from engine.risk import HedgingBacktest, StressTestEngine  # noqa: E402

PARAMS = {
    "svj_default": dict(kappa=3.0, theta=0.04, xi=0.5, rho=-0.7, v0=0.04, lambda_j=1.0, mu_j=-0.05, sigma_j=0.10, r=0.065, q=0.012),
    "gbm_cfg1": dict(kappa=3.0, theta=0.09, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=0.065, q=0.0),
}
STRESS = [("svj_default", 22500.0, 22500.0, 0.25, True, 2048), ("gbm_cfg1", 2500.0, 2600.0, 0.5, False, 1024)]
HEDGE = [("svj_default", 22500.0, 22500.0, 0.1, True, 9, 256, None, 5.0, 2.0),
         ("gbm_cfg1", 2500.0, 2450.0, 0.25, False, 25, 128, 20, 3.0, 1.0)]


def clean(x):
    if isinstance(x, dict):
        return {k: clean(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [clean(v) for v in x]
    return float(x) if not isinstance(x, (int, str)) else x


out = {"params": PARAMS, "stress": [], "hedge": []}
for name, spot, strike, T, call, n in STRESS:
    rep = StressTestEngine(SVJParams(**PARAMS[name]), num_paths=n, seed=42).full_stress_report(spot, strike, T, call)
    out["stress"].append(dict(params=name, spot=spot, strike=strike, T=T, is_call=call, num_paths=n, seed=42, report=clean(rep)))
for name, spot, strike, T, call, nsc, npaths, days, txn, slip in HEDGE:
    res = HedgingBacktest(SVJParams(**PARAMS[name]), seed=7).run_backtest(spot, strike, T, call, num_days=days, txn_cost_bps=txn,
                                                                        slippage_bps=slip, num_scenarios=nsc, num_mc_paths=npaths)
    out["hedge"].append(dict(params=name, spot=spot, strike=strike, T=T, is_call=call, num_scenarios=nsc, num_mc_paths=npaths,
                             num_days=days, txn_cost_bps=txn, slippage_bps=slip, seed=7, result=clean(res)))
with open(os.path.join(HERE, "risk_callers_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote risk_callers_golden.json:", len(out["stress"]), "stress reports,", len(out["hedge"]), "backtests")
