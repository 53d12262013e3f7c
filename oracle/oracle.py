"""
oracle/oracle.py -- CPU restatement of the reference's Monte Carlo hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker*.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the shipped package
(``monte_carlo_option_simulator_b200``) never does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  The reference is Python and imports in the dev container; every function
below is compared (tests/test_oracle.py) with fixtures written by running the reference itself
(tests/golden/make_golden.py, committed next to the fixtures it produced).

Each function cites the reference lines (relative to the reference tree) it restates.  The path
recurrence itself lives in oracle/svj_oracle.c (OpenMP C); ``simulate_svj`` calls it through ctypes
and ``simulate_svj_numpy`` is the same arithmetic in vectorised NumPy for boxes without the .so.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


# --------------------------------------------------------------------------------------------
# parameters (field set of engine/models.py:31-44)
# --------------------------------------------------------------------------------------------
@dataclass
class Params:
    kappa: float = 3.0
    theta: float = 0.04
    xi: float = 0.5
    rho: float = -0.7
    v0: float = 0.04
    lambda_j: float = 1.0
    mu_j: float = -0.05
    sigma_j: float = 0.10
    r: float = 0.065
    q: float = 0.012

    def replace(self, **kw) -> "Params":
        d = dict(self.__dict__)
        d.update(kw)
        return Params(**d)


def as_params(p) -> Params:
    """Accept the reference's SVJParams, the product's, or ours (duck-typed on field names)."""
    return Params(**{k: float(getattr(p, k)) for k in Params.__dataclass_fields__})


# --------------------------------------------------------------------------------------------
# C kernel loader
# --------------------------------------------------------------------------------------------
def build(force: bool = False) -> str:
    """Compile oracle/svj_oracle.c with the committed Makefile.  Returns the .so path."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "svj_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    lib = ctypes.CDLL(_LIB_PATH)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.oracle_simulate_svj.restype = None
    lib.oracle_simulate_svj.argtypes = [ctypes.c_double] * 12 + [dp, dp, dp, dp,
                                                                 ctypes.c_int64, ctypes.c_int32, ctypes.c_int,
                                                                 dp, dp, dp]
    u32p = ctypes.POINTER(ctypes.c_uint32)
    lib.oracle_philox4x32_10.restype = None
    lib.oracle_philox4x32_10.argtypes = [u32p, u32p, u32p]
    lib.oracle_philox_block_words.restype = None
    lib.oracle_philox_block_words.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64,
                                              ctypes.c_int32, ctypes.c_uint32, u32p]
    lib.oracle_num_threads.restype = ctypes.c_int
    u64p = ctypes.POINTER(ctypes.c_uint64)
    lib.oracle_np_standard_normal.restype = ctypes.c_uint64
    lib.oracle_np_standard_normal.argtypes = [u64p, u64p, dp, dp, ctypes.c_int64, dp]
    lib.oracle_glibc_log1p_fma.restype = ctypes.c_double
    lib.oracle_glibc_log1p_fma.argtypes = [ctypes.c_double]
    lib.oracle_log1p_mismatches.restype = ctypes.c_int64
    lib.oracle_log1p_mismatches.argtypes = [ctypes.c_int64, ctypes.c_uint64]
    _lib = lib
    return lib


def num_threads() -> int:
    return int(_load().oracle_num_threads())


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------------------------
# a1: engine/monte_carlo.py:189-243
# --------------------------------------------------------------------------------------------
def simulate_svj(S0, v0, r, q, T, kappa, theta, xi, rho, lambda_j, mu_j, sigma_j,
                 Z1, Z2, Z_jump, Z_jump_size, num_steps, record_paths=False):
    """Same signature and return triple as ``_simulate_svj_paths_numba`` (monte_carlo.py:190-199)."""
    lib = _load()
    Z1 = np.ascontiguousarray(Z1, dtype=np.float64)
    Z2 = np.ascontiguousarray(Z2, dtype=np.float64)
    Z_jump = np.ascontiguousarray(Z_jump, dtype=np.float64)
    Z_jump_size = np.ascontiguousarray(Z_jump_size, dtype=np.float64)
    n = Z1.shape[0]
    S = np.empty(n, dtype=np.float64)
    v = np.empty(n, dtype=np.float64)
    paths = np.empty((n, num_steps + 1), dtype=np.float64) if record_paths else np.zeros((0, 0))
    lib.oracle_simulate_svj(float(S0), v0, r, q, T, kappa, theta, xi, rho, lambda_j, mu_j, sigma_j,
                            _dptr(Z1), _dptr(Z2), _dptr(Z_jump), _dptr(Z_jump_size),
                            n, int(num_steps), int(bool(record_paths)),
                            _dptr(S), _dptr(v), _dptr(paths) if record_paths else None)
    return S, v, paths


def simulate_svj_numpy(S0, v0, r, q, T, kappa, theta, xi, rho, lambda_j, mu_j, sigma_j,
                       Z1, Z2, Z_jump, Z_jump_size, num_steps, record_paths=False):
    """monte_carlo.py:205-243 in vectorised NumPy (step loop outside, all paths at once)."""
    n = Z1.shape[0]
    dt = T / num_steps
    sqrt_dt = np.sqrt(dt)
    k = np.exp(mu_j + 0.5 * sigma_j ** 2) - 1.0
    drift_comp = r - q - lambda_j * k
    S = np.full(n, float(S0))
    v = np.full(n, float(v0))
    paths = np.zeros((n, num_steps + 1)) if record_paths else np.zeros((0, 0))
    if record_paths:
        paths[:, 0] = S0
    c2 = np.sqrt(1.0 - rho * rho)
    for s in range(num_steps):
        v_pos = np.maximum(v, 0.0)
        sqrt_v = np.sqrt(v_pos)
        dW1 = Z1[:, s] * sqrt_dt
        dW2 = rho * Z1[:, s] * sqrt_dt + c2 * Z2[:, s] * sqrt_dt
        jump = np.where(Z_jump[:, s] < lambda_j * dt, mu_j + sigma_j * Z_jump_size[:, s], 0.0)
        S = S * np.exp((drift_comp - 0.5 * v_pos) * dt + sqrt_v * dW1 + jump)
        v = np.maximum(v_pos + kappa * (theta - v_pos) * dt + xi * sqrt_v * dW2, 0.0)
        if record_paths:
            paths[:, s + 1] = S
    return S, v, paths


def _sim(p: Params, spot, T, Z1, Z2, Zj, Zjs, steps, record=False, v0=None):
    return simulate_svj(float(spot), p.v0 if v0 is None else v0, p.r, p.q, T, p.kappa, p.theta, p.xi, p.rho,
                        p.lambda_j, p.mu_j, p.sigma_j, Z1, Z2, Zj, Zjs, steps, record)


# --------------------------------------------------------------------------------------------
# a4: engine/monte_carlo.py:28-55
# --------------------------------------------------------------------------------------------
def _ncdf(x: float) -> float:
    from scipy.stats import norm
    return float(norm.cdf(x))


def bs_price(S, K, T, r, q, sigma, is_call=True) -> float:
    """monte_carlo.py:28-42."""
    if T <= 0:
        return max(S - K, 0.0) if is_call else max(K - S, 0.0)
    sT = sigma * np.sqrt(T)
    d1 = (np.log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / sT
    d2 = d1 - sT
    if is_call:
        return S * np.exp(-q * T) * _ncdf(d1) - K * np.exp(-r * T) * _ncdf(d2)
    return K * np.exp(-r * T) * _ncdf(-d2) - S * np.exp(-q * T) * _ncdf(-d1)


def bs_delta(S, K, T, r, q, sigma, is_call=True) -> float:
    """monte_carlo.py:45-55."""
    if T <= 0:
        if is_call:
            return 1.0 if S > K else 0.0
        return -1.0 if S < K else 0.0
    d1 = (np.log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / (sigma * np.sqrt(T))
    if is_call:
        return np.exp(-q * T) * _ncdf(d1)
    return np.exp(-q * T) * (_ncdf(d1) - 1.0)


# --------------------------------------------------------------------------------------------
# a2: pseudo-random front end, engine/monte_carlo.py:301-308 (same in :392-398, greeks.py:33-41)
# --------------------------------------------------------------------------------------------
def draw_pcg64(seed: int, n: int, steps: int):
    """Returns (Z1, Z2, Z_jump, Z_jump_size) in the reference's draw order."""
    g = np.random.default_rng(seed)
    Z1 = g.standard_normal((n, steps))
    Z2 = g.standard_normal((n, steps))
    Zjs = g.standard_normal((n, steps))
    Zj = np.random.default_rng(seed + 1).random((n, steps))
    return Z1, Z2, Zj, Zjs


def draw_sample_paths_rng(seed: int, n: int, steps: int):
    """monte_carlo.py:458-462: ONE generator at seed+999, order Z1, Z2, Z_jump_size, Z_jump."""
    g = np.random.default_rng(seed + 999)
    Z1 = g.standard_normal((n, steps))
    Z2 = g.standard_normal((n, steps))
    Zjs = g.standard_normal((n, steps))
    Zj = g.random((n, steps))
    return Z1, Z2, Zj, Zjs


# --------------------------------------------------------------------------------------------
# a2': quasi-random front end, engine/monte_carlo.py:61-183
# --------------------------------------------------------------------------------------------
def sobol_normals(num_paths: int, num_dims: int, seed: int = 0) -> np.ndarray:
    """monte_carlo.py:61-85: scrambled Sobol, 2^ceil(log2 n) points, clip, inverse normal CDF."""
    from scipy.stats import norm
    from scipy.stats.qmc import Sobol
    m = int(np.ceil(np.log2(max(num_paths, 2))))
    u = Sobol(d=num_dims, scramble=True, seed=seed).random(2 ** m)
    u = np.clip(u, 1e-10, 1 - 1e-10)
    return norm.ppf(u)[:num_paths]


def bb_order(n: int) -> List[int]:
    """monte_carlo.py:148-169: endpoint first, then breadth-first interval bisection."""
    if n <= 0:
        return []
    seen = [False] * n
    order = [n - 1]
    seen[n - 1] = True
    fifo = [(0, n - 1)]
    head = 0
    while head < len(fifo) and len(order) < n:
        lo, hi = fifo[head]
        head += 1
        if hi - lo <= 1:
            if not seen[lo] and len(order) < n:
                order.append(lo)
                seen[lo] = True
            continue
        mid = (lo + hi) // 2
        if not seen[mid]:
            order.append(mid)
            seen[mid] = True
        fifo.append((lo, mid))
        fifo.append((mid, hi))
    for i in range(n):
        if not seen[i]:
            order.append(i)
            seen[i] = True
    return order[:n]


def bb_reorder(normals: np.ndarray, num_steps: int) -> np.ndarray:
    """monte_carlo.py:88-145 incl. :172-183.  Reproduces the reference exactly, INCLUDING its quirk
    (SURVEY.md section 0, quirk 1): the first placed point is the endpoint itself, so t == t_right,
    var == 0 and W_T == 0 for every path."""
    n_paths = normals.shape[0]
    dt = 1.0 / num_steps
    order = bb_order(num_steps)
    W = np.zeros((n_paths, num_steps + 1))
    placed = np.zeros(num_steps + 2, dtype=bool)   # index = "actual" time index (idx + 1), :177
    for dim, tidx in enumerate(order):
        if dim >= normals.shape[1]:
            break
        tgt = tidx + 1
        t = tgt * dt
        # nearest placed neighbours, both INCLUSIVE of tgt itself (:179-182)
        left, right = 0, num_steps
        lo = np.flatnonzero(placed[1:tgt + 1])
        if lo.size:
            left = int(lo[-1]) + 1
        hi = np.flatnonzero(placed[tgt:num_steps])
        if hi.size:
            right = int(hi[0]) + tgt
        tl, tr = left * dt, right * dt
        if right > left:
            mu = W[:, left] + (W[:, right] - W[:, left]) * (t - tl) / (tr - tl)
            var = (t - tl) * (tr - t) / (tr - tl)
        else:
            mu = W[:, left]
            var = t - tl
        W[:, tgt] = mu + np.sqrt(max(var, 0)) * normals[:, dim]
        placed[tgt] = True
    return np.diff(W, axis=1)


def draw_sobol(seed: int, n: int, steps: int):
    """monte_carlo.py:290-299,306-308."""
    raw = sobol_normals(n, 3 * steps, seed=seed)
    Z1 = bb_reorder(raw[:, :steps], steps)
    Z2 = bb_reorder(raw[:, steps:2 * steps], steps)
    Zjs = np.ascontiguousarray(raw[:, 2 * steps:3 * steps])
    Zj = np.random.default_rng(seed + 1).random((n, steps))
    return Z1, Z2, Zj, Zjs


# --------------------------------------------------------------------------------------------
# a3 / a4 / a5: engine/monte_carlo.py:249-471
# --------------------------------------------------------------------------------------------
def _payoff(S, K, is_call):
    return np.maximum(S - K, 0.0) if is_call else np.maximum(K - S, 0.0)


def steps_for(num_steps: int, T: float, floor: int = 10) -> int:
    """monte_carlo.py:287 (floor 10) and :455 (floor 50)."""
    return max(int(num_steps * T), floor)


class MonteCarloOracle:
    """Restates MonteCarloEngine (monte_carlo.py:249-471) on top of the oracle kernel."""

    def __init__(self, params, num_paths=500_000, num_steps=252, seed=42,
                 use_sobol=True, use_antithetic=True, use_control_variate=True):
        self.p = as_params(params)
        self.n = int(num_paths)
        self.num_steps = int(num_steps)
        self.seed = int(seed)
        self.use_sobol = use_sobol
        self.use_antithetic = use_antithetic
        self.use_control_variate = use_control_variate

    def _draw(self, steps):
        return (draw_sobol if self.use_sobol else draw_pcg64)(self.seed, self.n, steps)

    def terminal(self, spot, T):
        steps = steps_for(self.num_steps, T)
        Z1, Z2, Zj, Zjs = self._draw(steps)
        S, _, _ = _sim(self.p, spot, T, Z1, Z2, Zj, Zjs, steps)
        S_anti = None
        if self.use_antithetic:                       # :318-324
            S_anti, _, _ = _sim(self.p, spot, T, -Z1, -Z2, Zj, -Zjs, steps)
        return steps, S, S_anti

    def price(self, spot, strike, T, is_call=True) -> Dict[str, float]:
        p, n = self.p, self.n
        steps, S, S_anti = self.terminal(spot, T)
        disc = np.exp(-p.r * T)                        # :327
        a = _payoff(S, strike, is_call)
        pay = 0.5 * (a + _payoff(S_anti, strike, is_call)) if self.use_antithetic else a   # :339
        raw = disc * np.mean(pay)                      # :342
        out = {"price": raw, "std_error": disc * np.std(pay) / np.sqrt(n),
               "num_paths_used": n, "num_steps": steps}
        if self.use_control_variate:                   # :353-373 (the "pseudo-CV", quirk 2)
            ref = bs_price(spot, strike, T, p.r, p.q, np.sqrt(p.v0), is_call)
            bs_mc = disc * np.mean(a)
            out["price"] = raw - (bs_mc - ref)
            out["bs_cv_adjustment"] = bs_mc - ref
            out["bs_ref"] = ref
            out["raw_mc_price"] = raw
            out["std_error"] = disc * np.std(pay - (a - ref / disc)) / np.sqrt(n)
        return out

    def price_batch(self, spot, strikes, T, is_call=True) -> List[Dict[str, float]]:
        p, n = self.p, self.n
        _, S, S_anti = self.terminal(spot, T)
        disc = np.exp(-p.r * T)
        res = []
        for K in strikes:                              # :420-448
            a = _payoff(S, K, is_call)
            pay = 0.5 * (a + _payoff(S_anti, K, is_call)) if self.use_antithetic else a
            raw = disc * np.mean(pay)
            row = {"strike": K, "price": raw, "std_error": disc * np.std(pay) / np.sqrt(n)}
            if self.use_control_variate:
                ref = bs_price(spot, K, T, p.r, p.q, np.sqrt(p.v0), is_call)
                row["price"] = raw - (disc * np.mean(a) - ref)
                row["bs_ref"] = ref
            res.append(row)
        return res

    def get_sample_paths(self, spot, T, num_samples=50) -> np.ndarray:
        steps = steps_for(self.num_steps, T, floor=50)              # :455
        Z1, Z2, Zj, Zjs = draw_sample_paths_rng(self.seed, num_samples, steps)
        return _sim(self.p, spot, T, Z1, Z2, Zj, Zjs, steps, record=True)[2]


# --------------------------------------------------------------------------------------------
# a6-a9: engine/greeks.py:20-263
# --------------------------------------------------------------------------------------------
class GreeksOracle:
    def __init__(self, params, num_paths=500_000, num_steps=252, seed=42):
        self.p = as_params(params)
        self.n = int(num_paths)
        self.num_steps = int(num_steps)
        self.seed = int(seed)

    def _setup(self, T):
        steps = steps_for(self.num_steps, T)
        return steps, np.exp(-self.p.r * T), draw_pcg64(self.seed, self.n, steps)   # greeks.py:33-41

    def delta(self, spot, strike, T, is_call=True, bump=0.01):
        steps, disc, (Z1, Z2, Zj, Zjs) = self._setup(T)
        S = _sim(self.p, spot, T, Z1, Z2, Zj, Zjs, steps)[0]
        if is_call:                                    # greeks.py:71-76
            pw = disc * np.mean((S > strike) * S / spot)
        else:
            pw = -disc * np.mean((S < strike) * S / spot)
        up = _sim(self.p, spot * (1 + bump), T, Z1, Z2, Zj, Zjs, steps)[0]
        dn = _sim(self.p, spot * (1 - bump), T, Z1, Z2, Zj, Zjs, steps)[0]
        fd = (disc * np.mean(_payoff(up, strike, is_call)) -
              disc * np.mean(_payoff(dn, strike, is_call))) / (2 * spot * bump)       # :89
        return {"pathwise": float(pw), "finite_diff": float(fd),
                "diff_pct": float(abs(pw - fd) / max(abs(fd), 1e-10) * 100)}

    def vega(self, spot, strike, T, is_call=True, bump=0.01):
        steps, disc, (Z1, Z2, Zj, Zjs) = self._setup(T)
        v_up = self.p.v0 + bump                        # greeks.py:124-125
        v_dn = max(self.p.v0 - bump, 0.001)
        up = _sim(self.p, spot, T, Z1, Z2, Zj, Zjs, steps, v0=v_up)[0]
        dn = _sim(self.p, spot, T, Z1, Z2, Zj, Zjs, steps, v0=v_dn)[0]
        fd = (disc * np.mean(_payoff(up, strike, is_call)) -
              disc * np.mean(_payoff(dn, strike, is_call))) / (v_up - v_dn)           # :156
        return {"fd_vega_v0": float(fd), "vega_per_vol_point": float(fd * 2 * np.sqrt(self.p.v0))}

    def gamma(self, spot, strike, T, is_call=True, bump=0.01):
        steps, disc, (Z1, Z2, Zj, Zjs) = self._setup(T)
        h = spot * bump                                # greeks.py:179
        pr = [disc * np.mean(_payoff(_sim(self.p, s0, T, Z1, Z2, Zj, Zjs, steps)[0], strike, is_call))
              for s0 in (spot, spot + h, spot - h)]
        return {"gamma": float((pr[1] - 2 * pr[0] + pr[2]) / h ** 2),
                "price_up": float(pr[1]), "price_base": float(pr[0]), "price_down": float(pr[2])}

    def theta(self, spot, strike, T, is_call=True, dt=1 / 252):
        eng = MonteCarloOracle(self.p, self.n, self.num_steps, self.seed)       # defaults: greeks.py:211-212
        p1 = eng.price(spot, strike, T, is_call)["price"]
        p2 = eng.price(spot, strike, max(T - dt, dt), is_call)["price"]
        th = -(p1 - p2) / dt
        return {"theta_daily": float(th), "theta_annual": float(th * 252)}

    def rho(self, spot, strike, T, is_call=True, bump=0.0001):
        h = bump                                        # greeks.py:231-246 (num_steps NOT forwarded)
        up = MonteCarloOracle(self.p.replace(r=self.p.r + h), self.n, seed=self.seed)
        dn = MonteCarloOracle(self.p.replace(r=max(self.p.r - h, 0)), self.n, seed=self.seed)
        val = (up.price(spot, strike, T, is_call)["price"] - dn.price(spot, strike, T, is_call)["price"]) / (2 * h)
        return {"rho": float(val), "rho_per_rate_point": float(val / 100)}


# --------------------------------------------------------------------------------------------
# a10: engine/risk.py:117-173
# --------------------------------------------------------------------------------------------
def hill(losses: np.ndarray, k: Optional[int] = None) -> float:
    """risk.py:158-173."""
    n = len(losses)
    if k is None:
        k = max(int(np.sqrt(n)), 10)
    k = min(k, n - 1)
    desc = np.sort(losses)[::-1]
    if desc[k] <= 0:
        return float("nan")
    s = np.sum(np.log(desc[:k] / desc[k]))
    return float(k / s) if s > 0 else float("nan")


def risk_metrics(returns: np.ndarray, confidence: float = 0.99) -> Dict[str, float]:
    """risk.py:117-155."""
    srt = np.sort(returns)
    n = len(srt)
    cut = int(n * (1 - confidence))
    var = -srt[cut] if cut < n else -srt[0]
    cvar = -np.mean(srt[:cut]) if cut > 0 else -srt[0]
    mean = np.mean(returns)
    std = np.std(returns)
    z = (returns - mean) / max(std, 1e-10)
    skew = float(np.mean(z ** 3))
    kurt = float(np.mean(z ** 4))
    losses = -srt[srt < 0]
    tail = hill(losses) if len(losses) > 20 else np.nan
    return {"var": float(var), "cvar": float(cvar), "skewness": skew, "kurtosis": kurt,
            "excess_kurtosis": kurt - 3.0, "tail_index": float(tail), "mean": float(mean), "std": float(std)}


# --------------------------------------------------------------------------------------------
# a11: the callers of engine/risk.py -- stress ladders (:33-111) and the delta-hedging backtest (:238-337)
# --------------------------------------------------------------------------------------------
SPOT_SHOCKS = [-0.08, -0.05, -0.02, 0.02, 0.05, 0.08]     # engine/config.py:134
VOL_SHOCKS = [-0.05, 0.05]                                # engine/config.py:135
JUMP_SCENARIO_SIZE = 0.04                                 # engine/config.py:136


def vol_shocked_params(p: Params, shock: float) -> Params:
    """risk.py:61-68."""
    return Params(kappa=p.kappa, theta=max(p.theta + shock ** 2, 0.001), xi=p.xi, rho=p.rho,
                  v0=max(p.v0 + 2 * np.sqrt(p.v0) * shock, 0.001), lambda_j=p.lambda_j, mu_j=p.mu_j,
                  sigma_j=p.sigma_j, r=p.r, q=p.q)


class StressOracle:
    """Restates StressTestEngine (risk.py:23-111).  `engine_kw` are extra MonteCarloOracle arguments (the reference
    uses the engine defaults: Sobol, antithetic, control variate)."""

    def __init__(self, params, num_paths=200_000, seed=42, **engine_kw):
        self.p, self.n, self.seed, self.kw = as_params(params), num_paths, seed, engine_kw

    def _engine(self, p=None):
        return MonteCarloOracle(p or self.p, num_paths=self.n, seed=self.seed, **self.kw)

    def spot_shock_ladder(self, spot, strike, T, is_call=True):
        eng = self._engine()
        base = eng.price(spot, strike, T, is_call)["price"]
        out = []
        for shock in SPOT_SHOCKS:                                          # :40-49
            s = spot * (1 + shock)
            pr = eng.price(s, strike, T, is_call)["price"]
            out.append({"shock_pct": shock * 100, "spot": s, "price": pr, "pnl": pr - base,
                        "pnl_pct": (pr - base) / max(base, 1e-6) * 100})
        return out

    def vol_shock_ladder(self, spot, strike, T, is_call=True):
        base = self._engine().price(spot, strike, T, is_call)["price"]
        out = []
        for shock in VOL_SHOCKS:                                           # :60-77
            sp = vol_shocked_params(self.p, shock)
            pr = self._engine(sp).price(spot, strike, T, is_call)["price"]
            out.append({"vol_shock": shock * 100, "v0": sp.v0, "price": pr, "pnl": pr - base})
        return out

    def jump_scenario(self, spot, strike, T, is_call=True, gap_size=JUMP_SCENARIO_SIZE):
        eng = self._engine()
        base = eng.price(spot, strike, T, is_call)["price"]
        dn = eng.price(spot * (1 - gap_size), strike, T, is_call)["price"]   # :89-90
        up = eng.price(spot * (1 + gap_size), strike, T, is_call)["price"]   # :93-94
        return {"base_price": base, "gap_down_price": dn, "gap_down_pnl": dn - base, "gap_up_price": up,
                "gap_up_pnl": up - base, "gap_size_pct": gap_size * 100}

    def full_stress_report(self, spot, strike, T, is_call=True):
        return {"spot_shocks": self.spot_shock_ladder(spot, strike, T, is_call),
                "vol_shocks": self.vol_shock_ladder(spot, strike, T, is_call),
                "jump_scenario": self.jump_scenario(spot, strike, T, is_call)}


def hedge_walk(p, spot, strike, T, is_call, num_days, txn_cost_bps, slippage_bps, premiums, Z):
    """The daily delta-hedging walk of HedgingBacktest.run_backtest (risk.py:278-316), vectorised over the
    scenarios (rows of Z: one standard normal per scenario and day, the order the reference draws them in).
    Returns (final_pnl[n], total_txn_cost[n])."""
    from scipy.stats import norm
    p = as_params(p)
    Z = np.asarray(Z, dtype=np.float64)
    n = Z.shape[0]
    dt = T / num_days                                                      # :258
    sigma = np.sqrt(p.v0)                                                  # :260
    S = np.full(n, float(spot))
    cash = np.array(premiums, dtype=np.float64).copy()                     # :273
    hedge = np.zeros(n)
    total_cost = np.zeros(n)
    t_rem = T
    for day in range(num_days):
        if t_rem <= 0:                                                     # :279-280
            break
        d1 = (np.log(S / strike) + (p.r - p.q + 0.5 * sigma ** 2) * t_rem) / (sigma * np.sqrt(t_rem))
        delta = np.exp(-p.q * t_rem) * norm.cdf(d1) if is_call else np.exp(-p.q * t_rem) * (norm.cdf(d1) - 1.0)
        trade = delta - hedge                                              # :286
        cost = np.abs(trade) * S * (txn_cost_bps + slippage_bps) / 10000   # :287
        total_cost += cost
        cash -= trade * S + cost                                           # :289
        hedge = delta
        S = S * np.exp((p.r - p.q - 0.5 * p.v0) * dt + np.sqrt(p.v0 * dt) * Z[:, day])   # :293-294
        t_rem -= dt                                                        # :308
    payoff = np.maximum(S - strike, 0) if is_call else np.maximum(strike - S, 0)
    return cash + hedge * S - payoff, total_cost                           # :316


class HedgingOracle:
    """Restates HedgingBacktest.run_backtest (risk.py:238-337): premium of every scenario from a fresh engine seeded
    seed + scenario (default engine flags), the walk above on default_rng(seed) normals drawn scenario by scenario."""

    def __init__(self, params, seed=42, **engine_kw):
        self.p, self.seed, self.kw = as_params(params), seed, engine_kw

    def run_backtest(self, spot, strike, T, is_call=True, num_days=None, txn_cost_bps=5.0, slippage_bps=2.0,
                     num_scenarios=1000, num_mc_paths=50_000):
        if num_days is None:
            num_days = max(int(T * 252), 1)                                # :255-256
        Z = np.random.default_rng(self.seed).standard_normal((num_scenarios, num_days))   # :262,292 (same stream)
        prem = [MonteCarloOracle(self.p, num_paths=num_mc_paths, seed=self.seed + s, **self.kw)
                .price(spot, strike, T, is_call)["price"] for s in range(num_scenarios)]  # :271-273
        pnl, cost = hedge_walk(self.p, spot, strike, T, is_call, num_days, txn_cost_bps, slippage_bps, prem, Z)
        pct = {f"{q}%": float(np.percentile(pnl, q)) for q in (1, 5, 25, 50, 75, 95, 99)}
        return {"mean_pnl": float(np.mean(pnl)), "std_pnl": float(np.std(pnl)), "pnl_percentiles": pct,
                "risk_metrics": risk_metrics(pnl, 0.99), "num_scenarios": num_scenarios,
                "total_txn_cost_avg": float(cost[-1]),                    # :336: the LAST scenario's total, not a mean
                "_pnl": pnl, "_cost": cost}


# --------------------------------------------------------------------------------------------
# 8(f)-3: implied volatility, engine/surface.py:22-126.  The root finder itself is SciPy's brentq (scipy.optimize, the
# reference's requirements.txt pins no version; 1.18.1 in this image) -- called here exactly as the reference calls it.
# --------------------------------------------------------------------------------------------
def bs_call_price(S, K, T, r, q, sigma):
    """surface.py:22-28."""
    from scipy.stats import norm
    if T <= 1e-10 or sigma <= 1e-10:
        return max(S * np.exp(-q * T) - K * np.exp(-r * T), 0.0)
    d1 = (np.log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / (sigma * np.sqrt(T))
    d2 = d1 - sigma * np.sqrt(T)
    return S * np.exp(-q * T) * norm.cdf(d1) - K * np.exp(-r * T) * norm.cdf(d2)


def bs_put_price(S, K, T, r, q, sigma):
    """surface.py:31-37."""
    from scipy.stats import norm
    if T <= 1e-10 or sigma <= 1e-10:
        return max(K * np.exp(-r * T) - S * np.exp(-q * T), 0.0)
    d1 = (np.log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / (sigma * np.sqrt(T))
    d2 = d1 - sigma * np.sqrt(T)
    return K * np.exp(-r * T) * norm.cdf(-d2) - S * np.exp(-q * T) * norm.cdf(-d1)


def implied_vol(price, S, K, T, r, q, is_call=True, lo=0.001, hi=5.0):
    """surface.py:48-66."""
    from scipy.optimize import brentq
    pricer = bs_call_price if is_call else bs_put_price
    try:
        f = lambda sigma: pricer(S, K, T, r, q, sigma) - price                    # noqa: E731
        if f(lo) * f(hi) > 0:
            return None
        return brentq(f, lo, hi, xtol=1e-8, maxiter=200)
    except (ValueError, RuntimeError):
        return None


def extract_iv_surface(spot, r, q, strikes, maturities, call_prices, put_prices, bid_ask_spreads=None,
                       max_spread_pct=0.10):
    """surface.py:69-126."""
    n_mat, n_k = call_prices.shape
    iv_call = np.full((n_mat, n_k), np.nan)
    iv_put = np.full((n_mat, n_k), np.nan)
    valid = np.ones((n_mat, n_k), dtype=bool)
    for i in range(n_mat):
        for j in range(n_k):
            if bid_ask_spreads is not None:                                       # :98-103
                mid = 0.5 * (call_prices[i, j] + put_prices[i, j])
                if mid > 0 and bid_ask_spreads[i, j] / mid > max_spread_pct:
                    valid[i, j] = False
                    continue
            c = implied_vol(call_prices[i, j], spot, strikes[j], maturities[i], r, q, True)
            p = implied_vol(put_prices[i, j], spot, strikes[j], maturities[i], r, q, False)
            if c is not None:
                iv_call[i, j] = c
            else:
                valid[i, j] = False
            if p is not None:
                iv_put[i, j] = p
            else:
                valid[i, j] = False
    return {"iv_call": iv_call, "iv_put": iv_put, "valid_mask": valid, "strikes": strikes, "maturities": maturities}


# --------------------------------------------------------------------------------------------
# 8(f)-4: the WORKING quasi-Monte Carlo front end of the CUDA path (not in the reference, whose bridge is degenerate --
# see bb_reorder above).  Sobol points are SciPy's own; the bridge is the textbook construction.
# --------------------------------------------------------------------------------------------
def qmc_bridge_nodes(n: int):
    """[(t, l, r, wl, wr, sd)] in construction order: the endpoint, then interval midpoints breadth first; unit time
    steps, W[0] = 0.  W[t] = wl W[l] + wr W[r] + sd z_k."""
    nodes = [(n, 0, 0, 0.0, 0.0, math.sqrt(n))]
    cur = [(0, n)]
    while cur:
        nxt = []
        for l, r in cur:
            if r - l <= 1:
                continue
            m = (l + r) // 2
            nodes.append((m, l, r, (r - m) / (r - l), (m - l) / (r - l), math.sqrt((m - l) * (r - m) / (r - l))))
            nxt += [(l, m), (m, r)]
        cur = nxt
    return nodes


def qmc_bridge(z: np.ndarray) -> np.ndarray:
    """z [n_paths, steps] in bridge order -> unit-variance step normals [n_paths, steps]."""
    n = z.shape[1]
    W = np.zeros((z.shape[0], n + 1))
    for k, (t, l, r, wl, wr, sd) in enumerate(qmc_bridge_nodes(n)):
        W[:, t] = wl * W[:, l] + wr * W[:, r] + sd * z[:, k]
    return np.diff(W, axis=1)


def qmc_draws(seed: int, n_paths: int, steps: int, n_blocks: int, path_offset: int = 0):
    """(Z1, Z2, Z_jump, Z_jump_size) of the device front end: scrambled Sobol (SciPy), clip, norm.ppf, blocks
    [Z1 bridged][Z2 bridged][jump sizes][jump uniforms]; blocks beyond n_blocks come back as neutral arrays."""
    from scipy.stats import norm
    from scipy.stats.qmc import Sobol
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)          # "balance properties require n to be a power of 2"
        pts = Sobol(d=n_blocks * steps, scramble=True, seed=seed).random(path_offset + n_paths)[path_offset:]
    u = np.clip(pts, 1e-10, 1 - 1e-10)
    blk = lambda b: u[:, b * steps:(b + 1) * steps]                               # noqa: E731
    Z1 = qmc_bridge(norm.ppf(blk(0)))
    Z2 = qmc_bridge(norm.ppf(blk(1))) if n_blocks >= 2 else np.zeros_like(Z1)
    Zjs = norm.ppf(blk(2)) if n_blocks >= 4 else np.zeros_like(Z1)
    Zj = blk(3).copy() if n_blocks >= 4 else np.ones_like(Z1)
    return Z1, Z2, Zj, Zjs


# --------------------------------------------------------------------------------------------
# NumPy's PCG64 + Ziggurat standard_normal restated (checker of csrc/np_normal.cu; engine/monte_carlo.py:301-304)
# --------------------------------------------------------------------------------------------
def np_standard_normal(seed, n: int, tables):
    """(normals, generator outputs consumed) of np.random.default_rng(seed).standard_normal(n), computed by the C
    restatement oracle_np_standard_normal from the generator's initial state and the Ziggurat tables (ki, wi, fi)."""
    st = np.random.default_rng(seed).bit_generator.state["state"]
    m = (1 << 64) - 1
    state = np.array([st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m], dtype=np.uint64)
    ki, wi, fi = (np.ascontiguousarray(t) for t in tables)
    out = np.empty(int(n), dtype=np.float64)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    used = _load().oracle_np_standard_normal(state.ctypes.data_as(u64p), ki.ctypes.data_as(u64p), _dptr(wi), _dptr(fi), int(n),
                                             _dptr(out))
    return out, int(used)


def log1p_mismatches(n: int, seed: int = 0) -> int:
    """Arguments (of 2n) on which the restated glibc FMA log1p differs from the host's log1p by even one bit."""
    return int(_load().oracle_log1p_mismatches(int(n), int(seed)))


# --------------------------------------------------------------------------------------------
# Philox4x32-10 (not in the reference; the generator of the CUDA path).  NumPy mirror for KATs and
# for checking the device's raw words bit for bit.
# --------------------------------------------------------------------------------------------
def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: uint32[..., 4], key: uint32[..., 2] -> uint32[..., 4]."""
    c = np.array(ctr, dtype=np.uint64, copy=True)
    k0 = np.array(key[..., 0], dtype=np.uint64, copy=True)
    k1 = np.array(key[..., 1], dtype=np.uint64, copy=True)
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    for _ in range(10):
        p0 = c0 * np.uint64(M0)
        p1 = c2 * np.uint64(M1)
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        c0, c1, c2, c3 = n0, p1 & np.uint64(MASK), n2, p0 & np.uint64(MASK)
        k0 = (k0 + np.uint64(W0)) & np.uint64(MASK)
        k1 = (k1 + np.uint64(W1)) & np.uint64(MASK)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def philox_block_words(seed: int, path_offset: int, n_paths: int, n_blocks: int, stream: int) -> np.ndarray:
    """uint32[n_paths, n_blocks, 4] with ctr = (path_lo, path_hi, block, stream), key = (seed_lo, seed_hi)."""
    lib = _load()
    out = np.empty((n_paths, n_blocks, 4), dtype=np.uint32)
    lib.oracle_philox_block_words(seed & (2 ** 64 - 1), path_offset, n_paths, n_blocks, stream,
                                  out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    return out
