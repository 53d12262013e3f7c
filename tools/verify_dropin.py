"""Drop-in check against a checkout of the reference (NOT part of the test suite: the GPU box has no reference).

    python tools/verify_dropin.py /path/to/Monte-Carlo-Option-Simulator

1. runs the reference's own verify.py unpatched (Numba CPU path) and keeps its prices;
2. patches the reference with patch_reference("engine") and runs verify.py again on the GPU -- it must print
   "ALL TESTS PASSED";
3. exercises the callers the north star names with the patched engine: the FastAPI handlers /api/price, /api/greeks,
   /api/smile, /api/stress (called as coroutines), StressTestEngine, one calibration objective evaluation, and
   compares patched and unpatched prices within Monte Carlo error.
"""
import asyncio
import contextlib
import io
import math
import os
import runpy
import sys
import time

ref = os.path.abspath(sys.argv[1])
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, ref)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_verify():
    buf = io.StringIO()
    t0 = time.time()
    with contextlib.redirect_stdout(buf):
        runpy.run_path(os.path.join(ref, "verify.py"), run_name="__main__")
    return buf.getvalue(), time.time() - t0


out_cpu, t_cpu = run_verify()
assert "ALL TESTS PASSED" in out_cpu
print(f"[unpatched] verify.py: ALL TESTS PASSED in {t_cpu:.1f} s")

import engine.app  # noqa: E402  (imports every engine module)
import engine.calibration  # noqa: E402
from engine.models import SVJParams  # noqa: E402
from engine.monte_carlo import MonteCarloEngine as RefEngine  # noqa: E402

svj = SVJParams()
ref_price = RefEngine(svj, num_paths=50_000, use_sobol=False, use_control_variate=False).price(22500.0, 22500.0, 0.25, True)
# the reference's DEFAULT configuration (Sobol + antithetic + control variate), unpatched, for the exact comparison below
t0 = time.time()
ref_default = RefEngine(svj, num_paths=50_000).price(22500.0, 22500.0, 0.25, True)
ref_default_put = RefEngine(svj, num_paths=50_000, seed=7).price(22500.0, 23000.0, 0.08, False)
t_ref_default = (time.time() - t0) / 2
# the reference's PSEUDO-random configuration (NumPy PCG64 Ziggurat normals), unpatched: price, Greeks, sample paths
t0 = time.time()
ref_pcg = RefEngine(svj, num_paths=50_000, use_sobol=False).price(22500.0, 22500.0, 0.25, True)
t_ref_pcg = time.time() - t0
from engine.greeks import GreeksEngine as RefGreeks  # noqa: E402
t0 = time.time()
_rg = RefGreeks(svj, num_paths=50_000)
ref_greeks = (_rg.delta(22500.0, 22500.0, 0.25, True), _rg.vega(22500.0, 22500.0, 0.25, True), _rg.gamma(22500.0, 22500.0, 0.25, True))
t_ref_greeks = time.time() - t0
ref_paths = RefEngine(svj, num_paths=50_000).get_sample_paths(22500.0, 0.25, 50)
from engine.risk import StressTestEngine as RefStress  # noqa: E402
ref_stress = RefStress(svj, num_paths=20_000).full_stress_report(22500.0, 22500.0, 0.08, True)

from monte_carlo_option_simulator_b200 import patch_reference  # noqa: E402
done = patch_reference("engine")
print(f"[patch] rebound {len(done)} names: {', '.join(done)}")

out_gpu, t_gpu = run_verify()
print(out_gpu)
assert "ALL TESTS PASSED" in out_gpu and "FAIL" not in out_gpu
print(f"[patched] verify.py: ALL TESTS PASSED in {t_gpu:.1f} s (unpatched {t_cpu:.1f} s)")

from engine.monte_carlo import MonteCarloEngine  # noqa: E402
assert MonteCarloEngine is not RefEngine
ours = MonteCarloEngine(svj, num_paths=2_000_000, use_sobol=False, use_control_variate=False).price(22500.0, 22500.0, 0.25, True)
z = abs(ours["price"] - ref_price["price"]) / math.hypot(ours["std_error"], ref_price["std_error"])
print(f"[price] reference {ref_price['price']:.3f} +- {ref_price['std_error']:.3f} (50k paths, CPU)   "
      f"patched {ours['price']:.3f} +- {ours['std_error']:.3f} (2M paths, GPU)   |z| = {z:.2f}")
assert z < 3.5

# rng="reference": the reference's own draws regenerated on the device -> the reference's numbers, exactly
t0 = time.time()
ours_default = MonteCarloEngine(svj, num_paths=50_000, rng="reference").price(22500.0, 22500.0, 0.25, True)
ours_default_put = MonteCarloEngine(svj, num_paths=50_000, seed=7, rng="reference").price(22500.0, 23000.0, 0.08, False)
t_ours_default = (time.time() - t0) / 2
for a, b in ((ref_default, ours_default), (ref_default_put, ours_default_put)):
    for k, v in a.items():
        assert abs(b[k] - v) <= 1e-9 * max(1.0, abs(v)), (k, v, b[k])
from engine.risk import StressTestEngine as OurStress  # noqa: E402
ours_stress = OurStress(svj, num_paths=20_000, rng="reference").full_stress_report(22500.0, 22500.0, 0.08, True)
for sec in ("spot_shocks", "vol_shocks"):
    for ra, rb in zip(ref_stress[sec], ours_stress[sec]):
        for k, v in ra.items():
            assert abs(rb[k] - v) <= 1e-8 * max(1.0, abs(v)), (sec, k, v, rb[k])
for k, v in ref_stress["jump_scenario"].items():
    assert abs(ours_stress["jump_scenario"][k] - v) <= 1e-8 * max(1.0, abs(v)), (k, v)
print(f"[exact] default-flag price() (Sobol + antithetic + CV, 50k paths): reference {ref_default['price']:.9f} "
      f"({t_ref_default:.2f} s per call, CPU)   patched rng=reference {ours_default['price']:.9f} ({t_ours_default * 1e3:.1f} ms per call)"
      f"   every key of two price() dicts and of a full stress report equal to 1e-9 / 1e-8")

# the pseudo-random configuration: NumPy's PCG64 Ziggurat normals regenerated on the device bit for bit (csrc/np_normal.cu)
from engine.greeks import GreeksEngine  # noqa: E402
t0 = time.time()
ours_pcg = MonteCarloEngine(svj, num_paths=50_000, use_sobol=False, rng="reference").price(22500.0, 22500.0, 0.25, True)
t_ours_pcg = time.time() - t0
for k, v in ref_pcg.items():
    assert abs(ours_pcg[k] - v) <= 1e-9 * max(1.0, abs(v)), (k, v, ours_pcg[k])
t0 = time.time()
_og = GreeksEngine(svj, num_paths=50_000, rng="reference")
ours_greeks = (_og.delta(22500.0, 22500.0, 0.25, True), _og.vega(22500.0, 22500.0, 0.25, True), _og.gamma(22500.0, 22500.0, 0.25, True))
t_ours_greeks = time.time() - t0
for a, b in zip(ref_greeks, ours_greeks):
    for k, v in a.items():
        assert abs(b[k] - v) <= 1e-8 * max(1.0, abs(v)), (k, v, b[k])
ours_paths = MonteCarloEngine(svj, num_paths=50_000, rng="reference").get_sample_paths(22500.0, 0.25, 50)
import numpy as _np  # noqa: E402
assert ours_paths.shape == ref_paths.shape and _np.allclose(ours_paths, ref_paths, rtol=1e-10, atol=0)
print(f"[exact] use_sobol=False price() (PCG64 Ziggurat draws, 50k x 63): reference {ref_pcg['price']:.9f} ({t_ref_pcg:.2f} s, CPU)   "
      f"patched rng=reference {ours_pcg['price']:.9f} ({t_ours_pcg * 1e3:.1f} ms)   every key equal to 1e-9; "
      f"GreeksEngine delta/vega/gamma: reference {t_ref_greeks:.2f} s, patched {t_ours_greeks * 1e3:.1f} ms, every key equal to 1e-8 "
      f"(delta {ours_greeks[0]['pathwise']:.9f} vs {ref_greeks[0]['pathwise']:.9f}); get_sample_paths 50 x {ref_paths.shape[1]} equal to 1e-10")

app = engine.app
calls = (("price", app.price_option, app.PriceRequest(spot=22500.0, strike=22500.0, T=0.08, num_paths=50_000)),
         ("greeks", app.compute_greeks, app.GreeksRequest(spot=22500.0, strike=22500.0, T=0.08, num_paths=50_000)),
         ("stress", app.run_stress, app.StressRequest(spot=22500.0, strike=22500.0, T=0.08, num_paths=50_000)),
         ("smile", app.generate_smile, app.SmileRequest(spot=22500.0, T=0.08)),
         ("hedge", app.run_hedge_backtest, app.HedgeRequest(spot=22500.0, strike=22500.0, T=0.04, num_scenarios=20)))
for name, handler, req in calls:
    t0 = time.time()
    res = asyncio.run(handler(req))
    keys = list(res.keys()) if isinstance(res, dict) else type(res).__name__
    print(f"[api] /api/{name}: {time.time() - t0:.2f} s, keys {keys}")

t0 = time.time()
import numpy as np  # noqa: E402
obj = engine.calibration._heston_objective(np.array([3.0, 0.04, 0.5, -0.7, 0.04]), 22500.0,
                                           np.array([22000.0, 22500.0, 23000.0]), 0.08,
                                           np.array([700.0, 420.0, 230.0]), np.array([1 / 3, 1 / 3, 1 / 3]), 0.065, 0.012,
                                           True, num_paths=20_000, num_steps=100)
print(f"[calibration] _heston_objective -> {obj} in {time.time() - t0:.2f} s")

# full two-stage calibration (the heaviest caller of the hot path): SciPy differential evolution drives the patched,
# per-candidate-batched objectives
from engine.monte_carlo import bs_price as _bs  # noqa: E402
ks = np.linspace(0.9, 1.1, 11) * 22500.0
mkt = np.array([_bs(22500.0, K, 0.08, 0.065, 0.012, 0.16 + 0.4 * (1 - K / 22500.0) ** 2 + 0.2 * max(1 - K / 22500.0, 0), True) for K in ks])
t0 = time.time()
cal = engine.calibration.CalibrationEngine().calibrate(22500.0, ks, 0.08, mkt, True, num_paths=20_000)
print(f"[calibration] CalibrationEngine.calibrate (two DE stages, 11 strikes, 20k paths): {time.time() - t0:.1f} s, "
      f"stage1 nit={cal['stage1_result']['nit']} err={cal['stage1_result']['error']:.4g}, "
      f"stage2 nit={cal['stage2_result']['nit']} err={cal['stage2_result']['error']:.4g}, v0={cal['params'].v0:.4f}")

from engine.risk import StressTestEngine  # noqa: E402
t0 = time.time()
rep = StressTestEngine(svj, num_paths=200_000).full_stress_report(22500.0, 22500.0, 0.08, True)
print(f"[stress] full_stress_report: {time.time() - t0:.2f} s, sections {list(rep.keys())}")

from engine.risk import HedgingBacktest  # noqa: E402
t0 = time.time()
bt = HedgingBacktest(svj).run_backtest(22500.0, 22500.0, 0.08, True)          # defaults: 1000 scenarios x 50k paths
print(f"[hedge] run_backtest (1000 scenarios x 50k paths): {time.time() - t0:.3f} s, mean_pnl {bt['mean_pnl']:.2f}, "
      f"std_pnl {bt['std_pnl']:.2f}, VaR99 {bt['risk_metrics']['var']:.2f}")

# opt-in: one launch per DE generation (vectorized / deferred updating)
patch_reference("engine", batch_population=True)
t0 = time.time()
cal2 = engine.calibration.CalibrationEngine().calibrate(22500.0, ks, 0.08, mkt, True, num_paths=20_000)
print(f"[calibration] population-batched calibrate: {time.time() - t0:.1f} s, "
      f"stage1 nit={cal2['stage1_result']['nit']} err={cal2['stage1_result']['error']:.4g}, "
      f"stage2 nit={cal2['stage2_result']['nit']} err={cal2['stage2_result']['error']:.4g}, v0={cal2['params'].v0:.4f}")
print("DROP-IN OK")
