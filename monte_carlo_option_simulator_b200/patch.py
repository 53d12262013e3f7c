"""patch_reference(): rebind every import-by-name site of the reference to the CUDA-backed mirrors.

The reference binds the kernel and the engine classes by name in several modules:
engine/greeks.py:16, engine/risk.py:16, engine/calibration.py:19, engine/app.py:26, verify.py:28.  After this call
the FastAPI app, calibration.py, risk.py's StressTestEngine / HedgingBacktest and verify.py run on libb200mc
without any edit to their sources."""
from __future__ import annotations

import importlib
import sys
from typing import List

from . import greeks as _g
from . import monte_carlo as _mc
from . import risk as _r
from . import surface as _s


def _batched_objectives(calib):
    """Same objectives as engine/calibration.py:53-135, with the per-strike loop of price() calls (:78-89, :119-130)
    replaced by ONE price_batch launch per candidate.  Numerically identical: the reference re-creates the same paths
    for every strike (same engine, same seed), which is exactly price_batch's "paths shared across strikes", and the
    two methods apply the same control-variate formula (monte_carlo.py:365 and :447).  REGULARIZATION, check_feller and
    SVJParams are taken from the patched module itself."""
    import numpy as np

    def _floats(a):
        """Plain floats: the objectives run 1e4-1e5 times per calibration on a handful of numbers, where NumPy scalars
        cost three times what floats do (the arithmetic is the same IEEE double either way)."""
        return a.tolist() if isinstance(a, np.ndarray) else [float(v) for v in a]

    def _sum_sq(engine, spot, strikes, T, market_prices, weights, is_call):
        try:
            rows = engine.price_batch(spot, _floats(strikes), T, is_call=is_call)
        except Exception:
            return float(len(strikes))                      # every strike "failed": +1.0 each (:88-89)
        return float(sum(w * (row["price"] - m) ** 2 for w, row, m in zip(_floats(weights), rows, _floats(market_prices))))

    def _heston_objective(x, spot, strikes, T, market_prices, weights, r, q, is_call, num_paths=100_000, num_steps=100):
        kappa, theta, xi, rho, v0 = _floats(x)
        feller_penalty = 0.0
        if not calib.check_feller(kappa, theta, xi):
            feller_penalty = 10.0 * (xi ** 2 - 2 * kappa * theta) ** 2
        params = calib.SVJParams(kappa=kappa, theta=theta, xi=xi, rho=rho, v0=v0, lambda_j=0.0, mu_j=0.0, sigma_j=0.01,
                                 r=r, q=q)
        engine = calib.MonteCarloEngine(params, num_paths=num_paths, num_steps=num_steps, use_sobol=True,
                                        use_antithetic=True, use_control_variate=True)
        total = _sum_sq(engine, spot, strikes, T, market_prices, weights, is_call)
        reg = calib.REGULARIZATION["xi"] * xi ** 2 + calib.REGULARIZATION["rho"] * rho ** 2
        return total + reg + feller_penalty

    def _svj_objective(x_jump, heston_params, spot, strikes, T, market_prices, weights, r, q, is_call,
                       num_paths=100_000, num_steps=100):
        lambda_j, mu_j, sigma_j = _floats(x_jump)
        kappa, theta, xi, rho, v0 = _floats(heston_params)
        params = calib.SVJParams(kappa=kappa, theta=theta, xi=xi, rho=rho, v0=v0, lambda_j=lambda_j, mu_j=mu_j,
                                 sigma_j=sigma_j, r=r, q=q)
        engine = calib.MonteCarloEngine(params, num_paths=num_paths, num_steps=num_steps, use_sobol=True,
                                        use_antithetic=True, use_control_variate=True)
        total = _sum_sq(engine, spot, strikes, T, market_prices, weights, is_call)
        return total + calib.REGULARIZATION["lambda_j"] * lambda_j ** 2

    # ---- population forms (second half): x is [n_params, S], the return value [S] ------------------
    def _population_sum_sq(cols, spot, strikes, T, market_prices, weights, is_call, num_paths, num_steps):
        S = max(np.size(v) for v in cols.values())
        try:
            engine = calib.MonteCarloEngine(calib.SVJParams(), num_paths=num_paths, num_steps=num_steps, use_sobol=True,
                                            use_antithetic=True, use_control_variate=True)
            prices = engine.price_population(cols, spot, np.asarray(strikes, dtype=float), T, is_call)
        except Exception:
            return np.full(S, float(len(strikes)))
        err = (prices - np.asarray(market_prices, dtype=float)[None, :]) ** 2 * np.asarray(weights, dtype=float)[None, :]
        bad = ~np.isfinite(err)                                  # a strike that "failed" counts 1.0 (:88-89)
        return np.where(bad, 1.0, err).sum(axis=1)

    def _heston_population(x, spot, strikes, T, market_prices, weights, r, q, is_call, num_paths=100_000, num_steps=100):
        x = np.asarray(x, dtype=float)
        if x.ndim == 1:
            return _heston_objective(x, spot, strikes, T, market_prices, weights, r, q, is_call, num_paths, num_steps)
        kappa, theta, xi, rho, v0 = x
        viol = xi ** 2 - 2 * kappa * theta
        feller_penalty = np.where(2.0 * kappa * theta > xi * xi, 0.0, 10.0 * viol ** 2)
        cols = dict(kappa=kappa, theta=theta, xi=xi, rho=rho, v0=v0, lambda_j=0.0, mu_j=0.0, sigma_j=0.01, r=r, q=q)
        total = _population_sum_sq(cols, spot, strikes, T, market_prices, weights, is_call, num_paths, num_steps)
        return total + calib.REGULARIZATION["xi"] * xi ** 2 + calib.REGULARIZATION["rho"] * rho ** 2 + feller_penalty

    def _svj_population(x_jump, heston_params, spot, strikes, T, market_prices, weights, r, q, is_call,
                        num_paths=100_000, num_steps=100):
        x_jump = np.asarray(x_jump, dtype=float)
        if x_jump.ndim == 1:
            return _svj_objective(x_jump, heston_params, spot, strikes, T, market_prices, weights, r, q, is_call,
                                  num_paths, num_steps)
        lambda_j, mu_j, sigma_j = x_jump
        kappa, theta, xi, rho, v0 = heston_params
        cols = dict(kappa=kappa, theta=theta, xi=xi, rho=rho, v0=v0, lambda_j=lambda_j, mu_j=mu_j, sigma_j=sigma_j, r=r, q=q)
        total = _population_sum_sq(cols, spot, strikes, T, market_prices, weights, is_call, num_paths, num_steps)
        return total + calib.REGULARIZATION["lambda_j"] * lambda_j ** 2

    _heston_objective.population = _heston_population
    _svj_objective.population = _svj_population
    return _heston_objective, _svj_objective


def _population_de(scipy_de):
    """differential_evolution wrapper for engine/calibration.py:195-226: objectives that carry a `.population` form are
    run with vectorized=True / updating='deferred' (SciPy stays the optimiser; one launch per generation instead of one
    per candidate).  Everything else goes to SciPy untouched."""
    def differential_evolution(func, bounds, args=(), **kw):
        pop = getattr(func, "population", None)
        if pop is None:
            return scipy_de(func, bounds, args=args, **kw)
        kw.pop("workers", None)
        kw["vectorized"], kw["updating"] = True, "deferred"
        return scipy_de(pop, bounds, args=args, **kw)
    differential_evolution._b200mc_inner = scipy_de
    return differential_evolution


def patch_reference(package: str = "engine", batch_calibration: bool = True, batch_scenarios: bool = True,
                    batch_population: bool = False, rng: str = None) -> List[str]:
    """Returns the list of 'module.attribute' names that were rebound.  batch_calibration: also replace the two
    calibration objectives by versions that price all strikes of a candidate in one launch.
    batch_scenarios: also replace StressTestEngine / HedgingBacktest / LiquidityStress (engine/risk.py:23-337) by the
    mirrors that price all scenarios of a report in one launch; with False the reference's own
    classes keep running, one price() call per scenario, on the rebound MonteCarloEngine.
    batch_population (opt-in, needs batch_calibration): also wrap engine.calibration.differential_evolution so that a
    whole DE generation is ONE launch (SciPy's vectorized=True, updating='deferred').  This changes the optimiser's
    update rule from SciPy's default 'immediate' to 'deferred', so the calibrated parameters differ from the reference's
    run (both are noisy optimisers of the same objective); hence not the default.
    rng: sets the default draw mode (the B200MC_RNG environment variable) for the engines the reference constructs."""
    if rng is not None:
        # default draw mode of every engine the reference's code constructs from now on (it never passes rng=):
        # "philox" device draws (statistical agreement), "reference" the reference's own draws (its numbers to ~1e-9;
        # at GPU speed for its default use_sobol=True flags), "sobol" the working quasi-Monte Carlo front end
        if rng not in ("philox", "reference", "sobol"):
            raise ValueError("rng must be 'philox', 'reference' or 'sobol'")
        import os
        os.environ["B200MC_RNG"] = rng
    done = []

    def rebind(modname, attr, obj):
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                return
        if hasattr(mod, attr):
            setattr(mod, attr, obj)
            done.append(f"{modname}.{attr}")

    mc = f"{package}.monte_carlo"
    for attr in ("_simulate_svj_paths_numba", "MonteCarloEngine", "bs_price", "bs_delta"):
        rebind(mc, attr, getattr(_mc, attr))
    rebind(f"{package}.greeks", "_simulate_svj_paths_numba", _mc._simulate_svj_paths_numba)
    rebind(f"{package}.greeks", "MonteCarloEngine", _mc.MonteCarloEngine)
    rebind(f"{package}.greeks", "GreeksEngine", _g.GreeksEngine)
    rebind(f"{package}.risk", "MonteCarloEngine", _mc.MonteCarloEngine)
    rebind(f"{package}.risk", "compute_risk_metrics", _r.compute_risk_metrics)
    if batch_scenarios:
        for attr in ("StressTestEngine", "LiquidityStress", "HedgingBacktest"):
            rebind(f"{package}.risk", attr, getattr(_r, attr))
    rebind(f"{package}.calibration", "MonteCarloEngine", _mc.MonteCarloEngine)
    calib = sys.modules.get(f"{package}.calibration")
    if batch_calibration and calib is not None and all(hasattr(calib, a) for a in
                                                      ("_heston_objective", "_svj_objective", "REGULARIZATION",
                                                       "check_feller", "SVJParams")):
        h_obj, s_obj = _batched_objectives(calib)
        rebind(f"{package}.calibration", "_heston_objective", h_obj)
        rebind(f"{package}.calibration", "_svj_objective", s_obj)
        if batch_population and hasattr(calib, "differential_evolution"):
            de = getattr(calib.differential_evolution, "_b200mc_inner", calib.differential_evolution)
            rebind(f"{package}.calibration", "differential_evolution", _population_de(de))
    for attr in ("implied_vol", "extract_iv_surface"):       # one launch per chain
        rebind(f"{package}.surface", attr, getattr(_s, attr))
    app = [("MonteCarloEngine", _mc.MonteCarloEngine), ("GreeksEngine", _g.GreeksEngine),
           ("compute_risk_metrics", _r.compute_risk_metrics)]
    if batch_scenarios:
        app += [(a, getattr(_r, a)) for a in ("StressTestEngine", "LiquidityStress", "HedgingBacktest")]
    for attr, obj in app:
        rebind(f"{package}.app", attr, obj)
    return done
