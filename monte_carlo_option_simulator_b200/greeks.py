"""Host mirror of engine/greeks.py (GreeksEngine, :20-263) on top of the fused CUDA kernel.

The reference obtains every Greek by re-simulating with regenerated identical draws: 3 kernel runs each for
delta, vega and gamma (:65-80,:111-147,:181-185).  Here ONE launch carries all bumps as extra per-path states over
the same Philox draws (common random numbers for free):
    spot bumps   S_T(S0 (1 +- b)) = (1 +- b) S_T exactly (the variance process does not see the spot)
    v0 bumps     two extra variance states (or, for constant variance, two extra terminal weights)
    rate bumps   S_T e^{(r' - r) T} exactly (the drift is linear in r)
theta and rho keep the reference's construction -- finite differences of MonteCarloEngine.price with that
engine's DEFAULT flags (:211-215,:242-246) -- and rho_crn / pathwise vega are
offered under new keys.  ``rng="reference"`` reproduces the reference draw for draw (PCG64 on the host).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import numpy as np

from . import _lib
from ._lib import FP64, GREEKS, SUMS_FIELDS, Bumps
from .models import copy_with
from .monte_carlo import DEFAULT_NUM_PATHS, MonteCarloEngine, _default_rng, steps_for

_COL = {n: i for i, n in enumerate(SUMS_FIELDS)}


class GreeksEngine:
    def __init__(self, params, num_paths: int = DEFAULT_NUM_PATHS, num_steps: int = 252, seed: int = 42, *,
                 rng: Optional[str] = None, precision: str = "fp32", handle=None, comm=None):
        self.params = params
        self.num_paths = num_paths
        self.num_steps = num_steps
        self.seed = seed
        self.rng = rng or _default_rng()
        self.precision = precision
        self._handle = handle
        self.comm = comm
        self._cache_key = None
        self._cache_row = None

    @property
    def handle(self):
        if self._handle is None:
            self._handle = _lib.default_handle()
        return self._handle

    # ---- one fused launch for every bump --------------------------------------------------------------------
    def _fused(self, spot, strike, T, is_call, spot_bump=0.01, v0_bump=0.01, r_bump=0.0001) -> np.ndarray:
        p = self.params
        key = (float(spot), float(strike), float(T), bool(is_call), spot_bump, v0_bump, r_bump,
               _lib._param_values(p), self.num_paths, self.num_steps, self.seed)
        if key == self._cache_key:
            return self._cache_row
        bumps = Bumps(spot_bump, p.v0 + v0_bump, max(p.v0 - v0_bump, 0.001),      # greeks.py:124-125
                      p.r + r_bump, max(p.r - r_bump, 0))                         # greeks.py:235,239
        flags = GREEKS | (FP64 if self.precision == "fp64" else 0)
        steps = steps_for(self.num_steps, T)
        n = int(self.num_paths)
        if self.comm is not None and self.comm.world > 1:
            from .dist import sharded_sums
            row = sharded_sums(self.handle, self.comm, p, float(spot), float(T), steps, n, self.seed, [float(strike)],
                               is_call, flags, bumps)[0]
        else:
            row = self.handle.price_european(p, float(spot), float(T), steps, n, self.seed, [float(strike)], is_call,
                                             flags, bumps)[0]
        row = row.tolist()                   # plain floats for the scalar algebra of delta / vega / gamma below
        self._cache_key, self._cache_row = key, row
        return row

    # ---- rng="reference" helpers (greeks.py:33-51) -----------------------------------------------------------
    def _generate_shared_randoms(self, steps: int):
        if hasattr(self.handle, "numpy_fill") and os.environ.get("B200MC_REFERENCE_PCG64", "device") == "device":
            # NumPy's draws generated on the device and kept there (bit for bit, csrc/np_normal.cu); cached while
            # (seed, n, steps) stay the same, which is how the reference obtains its common random numbers (:33-41)
            key = (self.seed, int(self.num_paths), int(steps), id(self.handle))
            cache = getattr(self, "_ref_draws", None)
            if cache is None or cache[0] != key:
                if cache is not None:
                    cache[1].close()
                self._ref_draws = cache = (key, _lib.ReferenceDraws(self.handle, self.seed, self.num_paths, steps))
            return (cache[1],)
        g = np.random.default_rng(self.seed)
        Z1 = g.standard_normal((self.num_paths, steps))
        Z2 = g.standard_normal((self.num_paths, steps))
        Zjs = g.standard_normal((self.num_paths, steps))
        Zj = np.random.default_rng(self.seed + 1).random((self.num_paths, steps))
        return Z1, Z2, Zj, Zjs

    def _simulate_with_randoms(self, spot, T, Z1, Z2=None, Z_jump=None, Z_jump_size=None, steps=None, params=None):
        if isinstance(Z1, _lib.ReferenceDraws):            # device-resident draws: _generate_shared_randoms returned (d,)
            if steps is None:
                steps = Z2
            S, v, _ = Z1.simulate(params or self.params, float(spot), T)
            return S, v, np.zeros((0, 0))
        S, v, _ = self.handle.simulate_given_normals(params or self.params, float(spot), T, Z1, Z2, Z_jump,
                                                     Z_jump_size, steps)
        return S, v, np.zeros((0, 0))

    @staticmethod
    def _pay(S, K, is_call):
        return np.maximum(S - K, 0) if is_call else np.maximum(K - S, 0)

    # ---- a6 --------------------------------------------------------------------------------------------------
    def delta(self, spot: float, strike: float, T: float, is_call: bool = True, bump: float = 0.01) -> Dict[str, float]:
        p = self.params
        discount = math.exp(-p.r * T)
        if self.rng == "reference":
            steps = steps_for(self.num_steps, T)
            Z = self._generate_shared_randoms(steps)
            S = self._simulate_with_randoms(spot, T, *Z, steps)[0]
            if is_call:
                pw = discount * np.mean((S > strike) * S / spot)                  # :71-73
            else:
                pw = -discount * np.mean((S < strike) * S / spot)                 # :74-76
            up = self._simulate_with_randoms(spot * (1 + bump), T, *Z, steps)[0]
            dn = self._simulate_with_randoms(spot * (1 - bump), T, *Z, steps)[0]
            pay_up = discount * np.mean(self._pay(up, strike, is_call))
            pay_dn = discount * np.mean(self._pay(dn, strike, is_call))
        else:
            row = self._fused(spot, strike, T, is_call, spot_bump=bump)
            n = row[_COL["n"]]
            pw = (1.0 if is_call else -1.0) * discount * row[_COL["sum_pw_delta"]] / n
            pay_up = discount * row[_COL["sum_spot_up"]] / n
            pay_dn = discount * row[_COL["sum_spot_dn"]] / n
        fd = (pay_up - pay_dn) / (2 * spot * bump)                                # :89
        return {"pathwise": float(pw), "finite_diff": float(fd),
                "diff_pct": float(abs(pw - fd) / max(abs(fd), 1e-10) * 100)}

    # ---- a7 --------------------------------------------------------------------------------------------------
    def vega(self, spot: float, strike: float, T: float, is_call: bool = True, bump: float = 0.01) -> Dict[str, float]:
        p = self.params
        discount = math.exp(-p.r * T)
        v0_up = p.v0 + bump                                                       # :124
        v0_down = max(p.v0 - bump, 0.001)                                         # :125
        extra = {}
        if self.rng == "reference":
            steps = steps_for(self.num_steps, T)
            Z = self._generate_shared_randoms(steps)
            up = self._simulate_with_randoms(spot, T, *Z, steps, params=copy_with(p, v0=v0_up))[0]
            dn = self._simulate_with_randoms(spot, T, *Z, steps, params=copy_with(p, v0=v0_down))[0]
            pay_up = discount * np.mean(self._pay(up, strike, is_call))
            pay_dn = discount * np.mean(self._pay(dn, strike, is_call))
        else:
            row = self._fused(spot, strike, T, is_call, v0_bump=bump)
            n = row[_COL["n"]]
            pay_up = discount * row[_COL["sum_v0_up"]] / n
            pay_dn = discount * row[_COL["sum_v0_dn"]] / n
            if p.xi == 0 and p.lambda_j <= 0 and (p.kappa == 0 or p.theta == p.v0):
                # NEW: pathwise dP/dsigma for frozen variance (Black-Scholes dynamics)
                extra["pathwise_vega_sigma"] = float((1.0 if is_call else -1.0) * discount * row[_COL["sum_pw_vega"]] / n)
        fd = (pay_up - pay_dn) / (v0_up - v0_down)                                # :156
        out = {"fd_vega_v0": float(fd), "vega_per_vol_point": float(fd * 2 * math.sqrt(p.v0))}   # :159-160
        out.update(extra)
        return out

    # ---- a8 --------------------------------------------------------------------------------------------------
    def gamma(self, spot: float, strike: float, T: float, is_call: bool = True, bump: float = 0.01) -> Dict[str, float]:
        p = self.params
        discount = math.exp(-p.r * T)
        h = spot * bump                                                           # :179
        if self.rng == "reference":
            steps = steps_for(self.num_steps, T)
            Z = self._generate_shared_randoms(steps)
            pr = [discount * np.mean(self._pay(self._simulate_with_randoms(s0, T, *Z, steps)[0], strike, is_call))
                  for s0 in (spot, spot + h, spot - h)]
            p_base, p_up, p_down = pr
        else:
            row = self._fused(spot, strike, T, is_call, spot_bump=bump)
            n = row[_COL["n"]]
            p_base = discount * row[_COL["sum_a"]] / n
            p_up = discount * row[_COL["sum_spot_up"]] / n
            p_down = discount * row[_COL["sum_spot_dn"]] / n
        g = (p_up - 2 * p_base + p_down) / (h ** 2)                               # :196
        return {"gamma": float(g), "price_up": float(p_up), "price_base": float(p_base), "price_down": float(p_down)}

    # ---- a9 --------------------------------------------------------------------------------------------------
    def _engine(self, params, **kw):
        return MonteCarloEngine(params, num_paths=self.num_paths, seed=self.seed, rng=self.rng,
                                precision=self.precision, handle=self._handle, comm=self.comm, **kw)

    def theta(self, spot: float, strike: float, T: float, is_call: bool = True, dt: float = 1 / 252) -> Dict[str, float]:
        engine = self._engine(self.params, num_steps=self.num_steps)              # default flags, :211-212
        p1 = engine.price(spot, strike, T, is_call)
        p2 = engine.price(spot, strike, max(T - dt, dt), is_call)                 # :214-215
        th = -(p1["price"] - p2["price"]) / dt
        return {"theta_daily": float(th), "theta_annual": float(th * 252)}

    def rho(self, spot: float, strike: float, T: float, is_call: bool = True, bump: float = 0.0001) -> Dict[str, float]:
        p = self.params
        h = bump
        e_up = self._engine(copy_with(p, r=p.r + h))                              # num_steps not forwarded, :242
        e_dn = self._engine(copy_with(p, r=max(p.r - h, 0)))
        val = (e_up.price(spot, strike, T, is_call)["price"] - e_dn.price(spot, strike, T, is_call)["price"]) / (2 * h)
        out = {"rho": float(val), "rho_per_rate_point": float(val / 100)}
        if self.rng != "reference":
            # NEW: common-random-number rho from the fused launch (discount each bumped payoff at its own rate)
            row = self._fused(spot, strike, T, is_call, r_bump=bump)
            n = row[_COL["n"]]
            r_up, r_dn = p.r + h, max(p.r - h, 0)
            out["rho_crn"] = float((math.exp(-r_up * T) * row[_COL["sum_r_up"]] / n -
                                    math.exp(-r_dn * T) * row[_COL["sum_r_dn"]] / n) / (r_up - r_dn))
        return out

    def all_greeks(self, spot: float, strike: float, T: float, is_call: bool = True) -> Dict[str, Dict]:
        """:254-263."""
        return {"delta": self.delta(spot, strike, T, is_call), "vega": self.vega(spot, strike, T, is_call),
                "gamma": self.gamma(spot, strike, T, is_call), "theta": self.theta(spot, strike, T, is_call),
                "rho": self.rho(spot, strike, T, is_call)}
