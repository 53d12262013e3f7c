"""Host mirror of engine/monte_carlo.py on top of libb200mc (CUDA, sm_100a).

Same names, argument meaning, result dictionaries and error behaviour as the reference module, so its callers
(engine/app.py:142-149,219-223, engine/greeks.py, engine/risk.py:37-93, engine/calibration.py:78-85, verify.py:31-51)
run unchanged after ``patch_reference()``:

    bs_price, bs_delta                     monte_carlo.py:28-55      closed forms (scalar, host)
    _simulate_svj_paths_numba              monte_carlo.py:189-243    -> b200mc_simulate_given_normals
    MonteCarloEngine.price                 monte_carlo.py:273-375    -> b200mc_price_european (fused)
    MonteCarloEngine.price_batch           monte_carlo.py:377-450    -> b200mc_price_european, all strikes in one launch
    MonteCarloEngine.get_sample_paths      monte_carlo.py:452-471    -> b200mc_generate_paths

Random numbers.  ``rng="philox"`` (default) draws on the device (Philox4x32-10 in registers, nothing materialised);
results agree with the reference statistically (within 3 standard errors), not draw for draw.  ``rng="reference"``
reproduces the reference's own host draws -- PCG64 or scrambled Sobol + its Brownian-bridge re-ordering
(monte_carlo.py:61-183,290-308) -- and runs the recurrence on the GPU over those arrays; results then agree with the
reference to ~1e-12.  That mode exists for parity and is as slow as the reference's RNG front end.
The environment variable B200MC_RNG overrides the default.

Documented divergences: an ``int`` spot is cast to float (the reference silently truncates every step to int64,
SURVEY.md section 0 quirk 3); the reference's degenerate Sobol/Brownian-bridge path and its pseudo control variate
(quirks 1, 2) are reproduced as they are, and a genuine control-variate estimate is returned under NEW keys
(``price_cv_spot``, ``std_error_cv_spot``) by the philox mode only.
"""
from __future__ import annotations

import functools
import math
import os
from typing import Dict, List, Optional

import numpy as np

from . import _lib
from ._lib import ANTITHETIC, FP64, SUMS_FIELDS

DEFAULT_NUM_PATHS = 500_000      # engine/config.py:23
DEFAULT_NUM_STEPS = 252          # engine/config.py:24

_COL = {n: i for i, n in enumerate(SUMS_FIELDS)}
_I_N, _I_A, _I_B, _I_AA, _I_BB, _I_AB = (_COL[k] for k in ("n", "sum_a", "sum_b", "sum_aa", "sum_bb", "sum_ab"))
_I_S, _I_SS, _I_PS = (_COL[k] for k in ("sum_s", "sum_ss", "sum_ps"))


# --------------------------------------------------------------------------------------------------------
# closed forms (engine/monte_carlo.py:28-55)
# --------------------------------------------------------------------------------------------------------
def _ncdf(x: float) -> float:
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def bs_price(S: float, K: float, T: float, r: float, q: float, sigma: float, is_call: bool = True) -> float:
    """Analytical Black-Scholes price (monte_carlo.py:28-42)."""
    if T <= 0:
        return max(S - K, 0.0) if is_call else max(K - S, 0.0)
    sT = sigma * math.sqrt(T)
    d1 = (math.log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / sT
    d2 = d1 - sT
    if is_call:
        return S * math.exp(-q * T) * _ncdf(d1) - K * math.exp(-r * T) * _ncdf(d2)
    return K * math.exp(-r * T) * _ncdf(-d2) - S * math.exp(-q * T) * _ncdf(-d1)


def bs_delta(S: float, K: float, T: float, r: float, q: float, sigma: float, is_call: bool = True) -> float:
    """Analytical Black-Scholes delta (monte_carlo.py:45-55)."""
    if T <= 0:
        if is_call:
            return 1.0 if S > K else 0.0
        return -1.0 if S < K else 0.0
    d1 = (math.log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / (sigma * math.sqrt(T))
    if is_call:
        return math.exp(-q * T) * _ncdf(d1)
    return math.exp(-q * T) * (_ncdf(d1) - 1.0)


# --------------------------------------------------------------------------------------------------------
# the reference's host RNG front end, kept on the host for rng="reference" (SURVEY.md 8a, rows a2 / a2')
# --------------------------------------------------------------------------------------------------------
def generate_sobol_normals(num_paths: int, num_dims: int, seed: int = 0) -> np.ndarray:
    """Scrambled Sobol points -> clip -> inverse normal CDF (monte_carlo.py:61-85)."""
    from scipy.stats import norm
    from scipy.stats.qmc import Sobol
    m = int(np.ceil(np.log2(max(num_paths, 2))))
    pts = Sobol(d=num_dims, scramble=True, seed=seed).random(2 ** m)
    return norm.ppf(np.clip(pts, 1e-10, 1 - 1e-10))[:num_paths]


def _bb_ordering(n: int) -> List[int]:
    """Visit order of the bridge: endpoint, then interval midpoints breadth first (monte_carlo.py:148-169)."""
    if n <= 0:
        return []
    done = np.zeros(n, dtype=bool)
    order = [n - 1]
    done[n - 1] = True
    work, head = [(0, n - 1)], 0
    while head < len(work) and len(order) < n:
        a, b = work[head]
        head += 1
        if b - a <= 1:
            if not done[a]:
                order.append(a)
                done[a] = True
            continue
        mid = (a + b) // 2
        if not done[mid]:
            order.append(mid)
            done[mid] = True
        work += [(a, mid), (mid, b)]
    order += [i for i in range(n) if not done[i]]
    return order[:n]


def brownian_bridge_reorder(normals: np.ndarray, num_steps: int) -> np.ndarray:
    """The reference's Brownian-bridge construction, monte_carlo.py:88-145 with the neighbour search of :172-183,
    reproduced INCLUDING its degeneracy (the first point placed is the endpoint itself, so its conditional variance
    is 0 and W_T == 0 on every path).  Kept for API parity only."""
    n_paths = normals.shape[0]
    dt = 1.0 / num_steps
    W = np.zeros((n_paths, num_steps + 1))
    known = np.zeros(num_steps + 2, dtype=bool)
    for dim, tidx in enumerate(_bb_ordering(num_steps)):
        if dim >= normals.shape[1]:
            break
        j = tidx + 1
        lo = np.flatnonzero(known[1:j + 1])
        hi = np.flatnonzero(known[j:num_steps])
        left = int(lo[-1]) + 1 if lo.size else 0
        right = int(hi[0]) + j if hi.size else num_steps
        t, tl, tr = j * dt, left * dt, right * dt
        if right > left:
            mean = W[:, left] + (W[:, right] - W[:, left]) * (t - tl) / (tr - tl)
            var = (t - tl) * (tr - t) / (tr - tl)
        else:
            mean = W[:, left]
            var = t - tl
        W[:, j] = mean + math.sqrt(max(var, 0)) * normals[:, dim]
        known[j] = True
    return np.diff(W, axis=1)


@functools.lru_cache(maxsize=64)
def reference_bridge_nodes(num_steps: int) -> np.ndarray:
    """The placements of brownian_bridge_reorder above as a b200mc_bridge_node table (_lib.BRIDGE_DTYPE, construction
    order): W[t] = W[l] + ((W[r] - W[l]) * a) / b + sd * z[dim], the reference's arithmetic (monte_carlo.py:128-133) with
    its neighbour search (:172-183).  A right neighbour that has not been placed yet is read as the zero it still holds
    (index 0) -- that is the first placement, the endpoint against itself, whose variance is therefore 0 (quirk 1)."""
    dt = 1.0 / num_steps
    known = np.zeros(num_steps + 2, dtype=bool)
    out = np.zeros(num_steps, dtype=_lib.BRIDGE_DTYPE)
    for dim, tidx in enumerate(_bb_ordering(num_steps)):
        j = tidx + 1
        lo = np.flatnonzero(known[1:j + 1])
        hi = np.flatnonzero(known[j:num_steps])
        left = int(lo[-1]) + 1 if lo.size else 0
        right = int(hi[0]) + j if hi.size else num_steps
        t, tl, tr = j * dt, left * dt, right * dt
        if right > left:
            a, b, var = t - tl, tr - tl, (t - tl) * (tr - t) / (tr - tl)
        else:
            a, b, var = 0.0, 1.0, t - tl
        r_idx = right if (right == 0 or known[right]) else 0
        out[dim] = (j, left, r_idx, dim, a, b, math.sqrt(max(var, 0)))
        known[j] = True
    return out


def _reference_draws(seed: int, n: int, steps: int, use_sobol: bool):
    """(Z1, Z2, Z_jump, Z_jump_size) exactly as MonteCarloEngine.price draws them (monte_carlo.py:290-308)."""
    if use_sobol:
        raw = generate_sobol_normals(n, 3 * steps, seed=seed)
        Z1 = brownian_bridge_reorder(raw[:, :steps], steps)
        Z2 = brownian_bridge_reorder(raw[:, steps:2 * steps], steps)
        Zjs = np.ascontiguousarray(raw[:, 2 * steps:3 * steps])
    else:
        g = np.random.default_rng(seed)
        Z1 = g.standard_normal((n, steps))
        Z2 = g.standard_normal((n, steps))
        Zjs = g.standard_normal((n, steps))
    Zj = np.random.default_rng(seed + 1).random((n, steps))
    return Z1, Z2, Zj, Zjs


# --------------------------------------------------------------------------------------------------------
# a1: drop-in for the Numba kernel
# --------------------------------------------------------------------------------------------------------
def _simulate_svj_paths_numba(S0, v0, r, q, T, kappa, theta, xi, rho, lambda_j, mu_j, sigma_j,
                              Z1, Z2, Z_jump, Z_jump_size, num_steps, record_paths=False):
    """Same signature and return triple as the reference kernel (monte_carlo.py:189-243); runs on the GPU."""
    class _P:
        pass
    p = _P()
    p.v0, p.r, p.q, p.kappa, p.theta, p.xi, p.rho = v0, r, q, kappa, theta, xi, rho
    p.lambda_j, p.mu_j, p.sigma_j = lambda_j, mu_j, sigma_j
    S, v, paths = _lib.default_handle().simulate_given_normals(p, float(S0), T, Z1, Z2, Z_jump, Z_jump_size,
                                                               int(num_steps), bool(record_paths))
    return S, v, (paths if record_paths else np.zeros((0, 0)))


def steps_for(num_steps: int, T: float, floor: int = 10) -> int:
    """monte_carlo.py:287 (floor 10) and :455 (floor 50)."""
    return max(int(num_steps * T), floor)


def _default_rng() -> str:
    return os.environ.get("B200MC_RNG", "philox")


# --------------------------------------------------------------------------------------------------------
# a3 / a4 / a5
# --------------------------------------------------------------------------------------------------------
class MonteCarloEngine:
    """Mirror of the reference's MonteCarloEngine (monte_carlo.py:249-471).

    Extra keyword-only arguments (not in the reference): ``rng`` ("philox" | "reference" | "sobol": a WORKING
    quasi-Monte Carlo front end -- SciPy's scrambled Sobol points generated on the device and a correct Brownian bridge,
    SURVEY.md 8f-4; the reference's own degenerate use_sobol path is what rng="reference" reproduces), ``precision``
    ("fp32" | "fp64": arithmetic of the path state in the fused kernels), ``handle`` (a ``_lib.Handle``),
    ``comm`` (a ``dist.Comm``: paths are sharded over its ranks and the sums all-reduced).
    """

    def __init__(self, params, num_paths: int = DEFAULT_NUM_PATHS, num_steps: int = DEFAULT_NUM_STEPS,
                 seed: int = 42, use_sobol: bool = True, use_antithetic: bool = True,
                 use_control_variate: bool = True, *, rng: Optional[str] = None, precision: str = "fp32",
                 handle=None, comm=None):
        self.params = params
        self.num_paths = num_paths
        self.num_steps = num_steps
        self.seed = seed
        self.use_sobol = use_sobol
        self.use_antithetic = use_antithetic
        self.use_control_variate = use_control_variate
        self.rng = rng or _default_rng()
        if self.rng not in ("philox", "reference", "sobol"):
            raise ValueError("rng must be 'philox', 'reference' or 'sobol'")
        if precision not in ("fp32", "fp64"):
            raise ValueError("precision must be 'fp32' or 'fp64'")
        self.precision = precision
        self._handle = handle
        self.comm = comm

    @property
    def handle(self):
        if self._handle is None:
            self._handle = _lib.default_handle()
        return self._handle

    def _extra_handles(self, k: int):
        """k more handles on the device of self.handle (own stream, scratch and counters), created once per engine."""
        have = getattr(self, "_lanes", [])
        while len(have) < k:
            have.append(_lib.Handle(self.handle.device))
        self._lanes = have
        return have[:max(k, 0)]

    # ---- fused path -------------------------------------------------------------------------------------
    def _flags(self) -> int:
        return (ANTITHETIC if self.use_antithetic else 0) | (FP64 if self.precision == "fp64" else 0)

    def _sobol(self, steps: int, T: float):
        """Scrambled direction numbers of SciPy's Sobol engine for the dimensions this run needs (cached)."""
        p = self.params
        nb = 4 if p.lambda_j * (T / steps) > 0.0 else (2 if p.xi != 0.0 else 1)
        key = (nb * steps, self.seed)
        cache = getattr(self, "_sobol_cache", None)
        if cache is None or cache[0] != key:
            self._sobol_cache = cache = (key, _lib.sobol_tables(nb * steps, self.seed))
        return cache[1]

    def _sums(self, spot, strikes, T, is_call, steps, flags=None, bumps=None) -> np.ndarray:
        flags = self._flags() if flags is None else flags
        n = int(self.num_paths)
        if self.rng == "sobol":
            if bumps is not None or (flags & _lib.GREEKS):
                raise ValueError("rng='sobol' has no Greek sums")
            lo, hi = 0, n
            world = self.comm.world if self.comm is not None else 1
            if world > 1:
                from .dist import shard_range
                lo, hi = shard_range(n, self.comm.rank, world)
            ks = np.atleast_1d(np.asarray(strikes, dtype=np.float64))
            rows = np.zeros((ks.size, len(SUMS_FIELDS)))
            if hi > lo:
                rows = self.handle.price_european_qmc(self.params, float(spot), float(T), steps, hi - lo, self._sobol(steps, T),
                                                      ks, is_call, flags & ANTITHETIC, path_offset=lo)
            return self.comm.allreduce_sum(rows).reshape(ks.size, -1) if world > 1 else rows
        if self.comm is not None and self.comm.world > 1:
            from .dist import sharded_sums
            return sharded_sums(self.handle, self.comm, self.params, float(spot), float(T), steps, n, self.seed,
                                strikes, is_call, flags, bumps)
        return self.handle.price_european(self.params, float(spot), float(T), steps, n, self.seed, strikes,
                                          is_call, flags, bumps)

    @staticmethod
    def _moments(row: np.ndarray, anti: bool):
        """mean and population variance of the combined payoff, mean of the primary payoff, variance of
        (combined - primary): everything price()/price_batch() need, from the five sums."""
        n, sa, sb, saa, sbb, sab = row[_I_N], row[_I_A], row[_I_B], row[_I_AA], row[_I_BB], row[_I_AB]
        if not n > 0:
            n = float("nan")                 # no paths: NaN results as the reference's empty reductions give, no exception
        mean_a = sa / n
        if anti:
            mean = 0.5 * (sa + sb) / n
            var = max(0.25 * (saa + 2.0 * sab + sbb) / n - mean * mean, 0.0)
            dmean = 0.5 * (sb - sa) / n
            dvar = max(0.25 * (saa - 2.0 * sab + sbb) / n - dmean * dmean, 0.0)
        else:
            mean, var = mean_a, max(saa / n - mean_a * mean_a, 0.0)
            dvar = 0.0
        return n, mean, var, mean_a, dvar

    def price(self, spot: float, strike: float, T: float, is_call: bool = True) -> Dict[str, float]:
        """Price a European option; same keys as monte_carlo.py:345-373."""
        if T == 0:
            return self._expired(spot, strike, is_call)
        if self.num_paths <= 0:              # the reference reduces empty arrays: NaN price and error, no exception
            res = {"price": float("nan"), "std_error": float("nan"), "num_paths_used": self.num_paths,
                   "num_steps": steps_for(self.num_steps, T)}
            if self.use_control_variate:
                res.update({"bs_cv_adjustment": float("nan"), "raw_mc_price": float("nan"),
                            "bs_ref": bs_price(float(spot), strike, T, self.params.r, self.params.q, math.sqrt(self.params.v0), is_call)})
            return res
        if self.rng == "reference":
            return self._price_reference(spot, strike, T, is_call)
        steps = steps_for(self.num_steps, T)                                   # :287
        row = self._sums(spot, [float(strike)], T, is_call, steps)[0]
        return self._result(row, self.params, spot, strike, T, is_call, steps)

    def _result(self, row, p, spot, strike, T, is_call, steps) -> Dict[str, float]:
        """The dict of price() from one b200mc_sums row."""
        if isinstance(row, np.ndarray):
            row = row.tolist()               # plain floats: the same IEEE arithmetic at a third of NumPy's scalar cost
        n, mean, var, mean_a, dvar = self._moments(row, self.use_antithetic)
        discount = math.exp(-p.r * T)                                          # :327
        raw_price = discount * mean                                            # :342
        result = {"price": raw_price, "std_error": discount * math.sqrt(var) / math.sqrt(n),   # :343
                  "num_paths_used": self.num_paths, "num_steps": steps}
        if self.use_control_variate:                                           # :353-373 (pseudo-CV, quirk 2)
            bs_ref = bs_price(float(spot), strike, T, p.r, p.q, math.sqrt(p.v0), is_call)
            bs_mc = discount * mean_a
            result["price"] = raw_price - (bs_mc - bs_ref)
            result["bs_cv_adjustment"] = bs_mc - bs_ref
            result["bs_ref"] = bs_ref
            result["raw_mc_price"] = raw_price
            result["std_error"] = discount * math.sqrt(dvar) / math.sqrt(n)
        result.update(self._spot_cv(row, spot, T, discount, p))
        return result

    def _expired(self, spot, strike, is_call, batch=False):
        """T == 0: the reference still walks its minimum of 10 steps with dt = 0 (monte_carlo.py:287), so every path ends at
        the spot and the result is the intrinsic value with zero standard error (its control-variate terms cancel:
        bs_price(T <= 0) is the intrinsic value too, :31-34).  No launch is needed for that.  (T < 0 has no meaning in the
        reference either -- sqrt(dt) is NaN there -- and raises here.)"""
        intrinsic = max(float(spot) - strike, 0.0) if is_call else max(strike - float(spot), 0.0)
        if batch:
            res = {"strike": strike, "price": intrinsic, "std_error": 0.0}
            if self.use_control_variate:
                res["bs_ref"] = intrinsic
            return res
        res = {"price": intrinsic, "std_error": 0.0, "num_paths_used": self.num_paths, "num_steps": steps_for(self.num_steps, 0.0)}
        if self.use_control_variate:
            res.update({"bs_cv_adjustment": 0.0, "bs_ref": intrinsic, "raw_mc_price": intrinsic})
        return res

    def price_many(self, spots, strikes, Ts, is_call=True, *, params=None, seeds=None) -> List[Dict[str, float]]:
        """NEW: many independent price() problems in ONE launch over a (cell x path) grid
        (b200mc_price_cells).  spots / strikes / Ts / is_call / params / seeds are scalars or sequences (broadcast);
        problem i returns exactly what MonteCarloEngine(params_i, num_paths, num_steps, seeds_i, <flags of self>)
        .price(spot_i, strike_i, T_i, is_call_i) returns (same draws, sums equal up to the order of the fp64 additions).
        This is what the loops of engine/risk.py:33-111 (stress ladders) and :264-273 (premium of every hedging
        scenario) become.  With a communicator the paths of every cell are sharded and the sums all-reduced once."""
        def seq(x):
            return list(x) if isinstance(x, (list, tuple, np.ndarray)) else None
        cols = [seq(spots), seq(strikes), seq(Ts), seq(is_call), seq(params), seq(seeds)]
        m = max([len(c) for c in cols if c is not None] or [1])
        defaults = [spots, strikes, Ts, is_call, params if params is not None else self.params,
                    seeds if seeds is not None else self.seed]
        cols = [c if c is not None else [d] * m for c, d in zip(cols, defaults)]
        if any(len(c) != m for c in cols):
            raise ValueError("price_many: sequences must have a common length")
        sp, ks, ts, calls, ps, sds = cols
        if self.rng != "philox":
            return [MonteCarloEngine(ps[i], self.num_paths, self.num_steps, sds[i], self.use_sobol, self.use_antithetic,
                                     self.use_control_variate, rng=self.rng, precision=self.precision,
                                     handle=self._handle, comm=self.comm).price(sp[i], ks[i], ts[i], calls[i])
                    for i in range(m)]
        steps = [steps_for(self.num_steps, float(T)) for T in ts]
        n, lo = int(self.num_paths), 0
        world = self.comm.world if self.comm is not None else 1
        if world > 1:
            from .dist import shard_range
            lo, hi = shard_range(n, self.comm.rank, world)
            n = hi - lo
        if n > 0:
            cells = _lib.make_cells(ps if isinstance(params, (list, tuple)) else ps[0], [float(x) for x in sp],
                                    [float(x) for x in ts], steps, n, sds, lo, [bool(c) for c in calls])
            sums = self.handle.price_cells(cells, np.asarray(ks, dtype=np.float64), self._flags())[:, 0, :]
        else:
            sums = np.zeros((m, len(SUMS_FIELDS)))
        if world > 1:
            sums = self.comm.allreduce_sum(sums).reshape(m, len(SUMS_FIELDS))
        return [self._result(sums[i], ps[i], sp[i], ks[i], ts[i], calls[i], steps[i]) for i in range(m)]

    def prices_for_seeds(self, spot: float, strike: float, T: float, is_call: bool, seeds) -> np.ndarray:
        """price()["price"] of this engine re-seeded with every entry of `seeds`, as one launch and one vectorised
        formula (the premiums of HedgingBacktest's scenarios, engine/risk.py:271-273)."""
        seeds = list(seeds)
        if self.rng != "philox" or (self.comm is not None and self.comm.world > 1):
            return np.array([r["price"] for r in self.price_many(spot, strike, T, is_call, seeds=seeds)])
        p = self.params
        steps = steps_for(self.num_steps, T)
        cells = _lib.make_cells(p, float(spot), float(T), steps, int(self.num_paths), seeds, 0, bool(is_call))
        sums = self.handle.price_cells(cells, np.full(len(seeds), float(strike)), self._flags())[:, 0, :]
        n, sa, sb = sums[:, _COL["n"]], sums[:, _COL["sum_a"]], sums[:, _COL["sum_b"]]
        discount = math.exp(-p.r * T)
        price = discount * (0.5 * (sa + sb) if self.use_antithetic else sa) / n                     # :342
        if self.use_control_variate:                                                                # :353-365
            bs_ref = bs_price(float(spot), strike, T, p.r, p.q, math.sqrt(p.v0), is_call)
            price = price - (discount * sa / n - bs_ref)
        return price

    def price_population(self, param_cols: Dict[str, np.ndarray], spot: float, strikes, T: float,
                         is_call: bool = True) -> np.ndarray:
        """NEW: price_batch for a whole POPULATION of parameter sets in one launch.  param_cols maps
        the ten SVJParams fields to scalars or arrays [S]; returns the prices [S, n_strikes] -- row s is what
        MonteCarloEngine(params_s, <this engine's settings>).price_batch(spot, strikes, T, is_call) reports as "price"
        (paths shared across the strikes of a candidate, same seed for every candidate as in the reference's
        objectives, engine/calibration.py:78-85).  Philox draws only."""
        if self.rng != "philox":
            raise ValueError("price_population draws on the device (rng='philox')")
        ks = np.ascontiguousarray(np.asarray(strikes, dtype=np.float64).ravel())
        steps = steps_for(self.num_steps, T)
        cells = _lib.make_cells(param_cols, float(spot), float(T), steps, int(self.num_paths), self.seed, 0, bool(is_call))
        S = cells.size
        sums = self.handle.price_cells(cells, np.tile(ks, (S, 1)), self._flags())             # [S, K, NSUMS]
        n = sums[:, :, _COL["n"]]
        sa, sb = sums[:, :, _COL["sum_a"]], sums[:, :, _COL["sum_b"]]
        r, q, v0 = cells["r"][:, None], cells["q"][:, None], cells["v0"][:, None]
        discount = np.exp(-r * T)
        mean = (0.5 * (sa + sb) if self.use_antithetic else sa) / n
        price = discount * mean
        if self.use_control_variate:                                            # monte_carlo.py:443-448
            from scipy.special import ndtr
            sig = np.sqrt(v0)
            sT = sig * math.sqrt(T)
            d1 = (np.log(float(spot) / ks[None, :]) + (r - q + 0.5 * sig ** 2) * T) / sT
            d2 = d1 - sT
            if is_call:
                bs_ref = float(spot) * np.exp(-q * T) * ndtr(d1) - ks[None, :] * discount * ndtr(d2)
            else:
                bs_ref = ks[None, :] * discount * ndtr(-d2) - float(spot) * np.exp(-q * T) * ndtr(-d1)
            price = price - (discount * sa / n - bs_ref)
        return price

    def _spot_cv(self, row, spot, T, discount, p=None) -> Dict[str, float]:
        """NEW keys: regression control variate on S_T, whose mean S0 e^{(r-q)T} is known for every SVJ
        parameter set (the jump drift is compensated, monte_carlo.py:209-210)."""
        p = p or self.params
        n = row[_I_N]
        if not n > 0:
            return {}
        anti = self.use_antithetic
        pay_mean = (0.5 * (row[_I_A] + row[_I_B]) if anti else row[_I_A]) / n
        if anti:
            pay_sq = 0.25 * (row[_I_AA] + 2 * row[_I_AB] + row[_I_BB]) / n
        else:
            pay_sq = row[_I_AA] / n
        s_mean, s_sq, ps = row[_I_S] / n, row[_I_SS] / n, row[_I_PS] / n
        var_s = s_sq - s_mean * s_mean
        if not var_s > 0:
            return {}
        cov = ps - pay_mean * s_mean
        beta = cov / var_s
        target = float(spot) * math.exp((p.r - p.q) * T)
        var_cv = max(pay_sq - pay_mean * pay_mean - cov * cov / var_s, 0.0)
        return {"price_cv_spot": discount * (pay_mean - beta * (s_mean - target)),
                "std_error_cv_spot": discount * math.sqrt(var_cv) / math.sqrt(n)}

    def price_batch(self, spot: float, strikes, T: float, is_call: bool = True) -> list:
        """Price multiple strikes with shared path simulation (monte_carlo.py:377-450)."""
        if T == 0:
            return [self._expired(spot, K, is_call, batch=True) for K in strikes]
        if self.rng == "reference":
            return self._price_batch_reference(spot, strikes, T, is_call)
        p = self.params
        steps = steps_for(self.num_steps, T)
        strikes = np.asarray(strikes).ravel().tolist() if isinstance(strikes, np.ndarray) else list(strikes)
        discount = math.exp(-p.r * T)
        sigma_bs = math.sqrt(p.v0)
        results = []
        # The per-strike algebra is scalar Python (a smile has 11-64 strikes; for the small calls of a calibration this
        # loop costs as much as the launch), so everything that does not depend on the strike is hoisted: the Black-
        # Scholes reference below is bs_price() term for term -- the same operations in the same order, bit-identical.
        anti, cv, moments, sqrt = self.use_antithetic, self.use_control_variate, self._moments, math.sqrt
        S = float(spot)
        if cv and T > 0:
            sT = sigma_bs * sqrt(T)
            drift = (p.r - p.q + 0.5 * sigma_bs ** 2) * T
            S_eq, e_r = S * math.exp(-p.q * T), math.exp(-p.r * T)
        for lo in range(0, len(strikes), 256):                                 # at most 256 strikes per launch
            chunk = strikes[lo:lo + 256]
            rows = self._sums(spot, chunk, T, is_call, steps)
            for K, row in zip(chunk, rows.tolist()):
                n, mean, var, mean_a, _ = moments(row, anti)
                raw = discount * mean
                res = {"strike": K, "price": raw, "std_error": discount * sqrt(var) / sqrt(n)}             # :438-441
                if cv:                                                         # :443-448
                    Kf = float(K)
                    if T > 0:
                        d1 = (math.log(S / Kf) + drift) / sT
                        d2 = d1 - sT
                        bs_ref = S_eq * _ncdf(d1) - Kf * e_r * _ncdf(d2) if is_call else \
                            Kf * e_r * _ncdf(-d2) - S_eq * _ncdf(-d1)
                    else:
                        bs_ref = bs_price(S, Kf, T, p.r, p.q, sigma_bs, is_call)
                    res["price"] = raw - (discount * mean_a - bs_ref)
                    res["bs_ref"] = bs_ref
                results.append(res)
        return results

    def price_grid(self, spot: float, strikes, maturities, is_call: bool = True, *,
                   independent_cells: bool = False) -> Dict[str, np.ndarray]:
        """NEW (BASELINE config 3): strike x maturity grid in the (n_mat, n_k) layout engine/surface.py:69-126
        (extract_iv_surface) consumes.  Returns prices, std_errors (raw, as price_batch), num_steps.

        independent_cells=False (API-parity reading): per expiry this is exactly price_batch -- paths shared across
        strikes, the steps rule of monte_carlo.py:287, same seed -- one fused launch per expiry; with a communicator the
        PATHS of every expiry are sharded and the sums all-reduced once.
        independent_cells=True (benchmark reading): every (expiry, strike) cell gets its own num_paths paths (disjoint
        Philox counter ranges: path_offset = cell * num_paths), one launch per cell; with a communicator the CELLS are
        dealt round-robin to the ranks and no path-level collective is needed (only the final gather of the sums).
        All launches are queued on the stream before one synchronisation and one device->host copy."""
        if self.rng != "philox":
            rows = [self.price_batch(spot, strikes, float(T), is_call) for T in maturities]
            steps = [steps_for(self.num_steps, float(T)) for T in maturities]
        else:
            from ._lib import NSUMS
            p = self.params
            ks = np.ascontiguousarray(np.asarray(strikes, dtype=np.float64).ravel())
            Ts = [float(T) for T in maturities]
            if ks.size > 256:
                raise ValueError("price_grid takes at most 256 strikes")
            h = self.handle
            steps = [steps_for(self.num_steps, T) for T in Ts]
            n = int(self.num_paths)
            rank, world = (self.comm.rank, self.comm.world) if self.comm is not None else (0, 1)
            nbytes = len(Ts) * ks.size * NSUMS * 8
            buf = h.malloc(nbytes)
            try:
                sums = np.zeros((len(Ts), ks.size, NSUMS))
                h.h2d(buf, sums)
                if independent_cells:
                    # cells are small launches (1M paths x 31..500 steps = 20..300 us): a few handles (= streams with
                    # their own scratch) on the same device let the tail of one launch overlap the head of the next
                    lanes = [h] + self._extra_handles(int(os.environ.get("B200MC_GRID_STREAMS", "4")) - 1)
                    mine = [(j, i) for j in range(len(Ts)) for i in range(ks.size) if (j * ks.size + i) % world == rank]
                    mine.sort(key=lambda ji: -steps[ji[0]])                  # longest first
                    h.synchronize()                                          # the zero-fill above is on h's stream
                    for q, (j, i) in enumerate(mine):
                        cell = j * ks.size + i
                        lanes[q % len(lanes)].price_european(p, float(spot), Ts[j], steps[j], n, self.seed, ks[i:i + 1],
                                                             is_call, self._flags(), None, path_offset=cell * n,
                                                             out_dev=buf + cell * NSUMS * 8)
                    for lane in lanes[1:]:
                        lane.synchronize()
                    h.d2h(sums, buf)
                else:
                    lo, hi = 0, n
                    if world > 1:
                        from .dist import shard_range
                        lo, hi = shard_range(n, rank, world)
                    if hi > lo:
                        for j, (T, st) in enumerate(zip(Ts, steps)):
                            h.price_european(p, float(spot), T, st, hi - lo, self.seed, ks, is_call, self._flags(), None,
                                             path_offset=lo, out_dev=buf + j * ks.size * NSUMS * 8)
                        h.d2h(sums, buf)
            finally:
                h.free(buf)
            if world > 1:
                sums = self.comm.allreduce_sum(sums).reshape(len(Ts), ks.size, NSUMS)
            rows = []
            for T, block in zip(Ts, sums):
                discount = math.exp(-p.r * T)
                out = []
                for K, row in zip(ks, block.tolist()):
                    nn, mean, var, mean_a, _ = self._moments(row, self.use_antithetic)
                    res = {"strike": K, "price": discount * mean, "std_error": discount * math.sqrt(var) / math.sqrt(nn)}
                    if self.use_control_variate:
                        bs_ref = bs_price(float(spot), float(K), T, p.r, p.q, math.sqrt(p.v0), is_call)
                        res["price"] = res["price"] - (discount * mean_a - bs_ref)
                    out.append(res)
                rows.append(out)
        return {"strikes": np.asarray(strikes, dtype=np.float64), "maturities": np.asarray(maturities, dtype=np.float64),
                "prices": np.array([[r["price"] for r in row] for row in rows]),
                "std_errors": np.array([[r["std_error"] for r in row] for row in rows]),
                "num_steps": np.asarray(steps)}

    def get_sample_paths(self, spot: float, T: float, num_samples: int = 50) -> np.ndarray:
        """[num_samples, steps + 1] float64, column 0 = spot (monte_carlo.py:452-471)."""
        steps = steps_for(self.num_steps, T, floor=50)                         # :455
        if self.rng == "reference" and hasattr(self.handle, "numpy_fill") and \
                os.environ.get("B200MC_REFERENCE_PCG64", "device") == "device":
            # :458-462 -- ONE generator (seed + 999): three normal arrays, then the jump uniforms from the same stream
            d = _lib.ReferenceDraws(self.handle, self.seed + 999, int(num_samples), steps, uniform_seed=None)
            try:
                return d.simulate(self.params, float(spot), T, record_paths=True)[2]
            finally:
                d.close()
        if self.rng == "reference":
            g = np.random.default_rng(self.seed + 999)                         # :458-462
            Z1 = g.standard_normal((num_samples, steps))
            Z2 = g.standard_normal((num_samples, steps))
            Zjs = g.standard_normal((num_samples, steps))
            Zj = g.random((num_samples, steps))
            return self.handle.simulate_given_normals(self.params, float(spot), T, Z1, Z2, Zj, Zjs, steps, True)[2]
        return self.handle.generate_paths(self.params, float(spot), T, steps, int(num_samples), self.seed + 999,
                                          FP64, np.float64)

    # ---- rng="reference": the reference's own draws, recurrence on the GPU ------------------------------
    def _terminal_reference(self, spot, T):
        p = self.params
        n = int(self.num_paths)
        steps = steps_for(self.num_steps, T)
        h = self.handle
        if self.use_sobol and hasattr(h, "qmc_terminal") and os.environ.get("B200MC_REFERENCE_SOBOL", "device") == "device":
            # the reference's Sobol front end (:290-299,308) on the device: SciPy's scrambled points (bitwise), norm.ppf,
            # the reference's own bridge table, NumPy's PCG64 uniforms for the jumps -- nothing is drawn on the host
            key = (steps, self.seed)
            cache = getattr(self, "_ref_sobol", None)
            if cache is None or cache[0] != key:
                self._ref_sobol = cache = (key, _lib.sobol_tables(3 * steps, self.seed), reference_bridge_nodes(steps))
            # (the jump uniforms of :308, default_rng(seed + 1).random((n, steps)), are reproduced on the device too)
            S, S_anti = h.qmc_terminal(p, float(spot), T, steps, n, cache[1], cache[2], None,
                                       ANTITHETIC if self.use_antithetic else 0, pcg64_seed=self.seed + 1)
            return steps, S, S_anti
        if not self.use_sobol and hasattr(h, "numpy_fill") and os.environ.get("B200MC_REFERENCE_PCG64", "device") == "device":
            # the reference's pseudo-random front end (:301-308) on the device: NumPy's PCG64 Ziggurat normals and uniforms
            # bit for bit (csrc/np_normal.cu), kept in HBM and reused while (seed, n, steps) stay the same
            d = self._reference_draws_device(n, steps)
            S = d.simulate(p, float(spot), T)[0]
            S_anti = d.simulate(p, float(spot), T, negate=True)[0] if self.use_antithetic else None      # :318-324
            return steps, S, S_anti
        Z1, Z2, Zj, Zjs = _reference_draws(self.seed, n, steps, self.use_sobol)
        S = h.simulate_given_normals(p, float(spot), T, Z1, Z2, Zj, Zjs, steps)[0]
        S_anti = None
        if self.use_antithetic:                                                # :318-324
            S_anti = h.simulate_given_normals(p, float(spot), T, -Z1, -Z2, Zj, -Zjs, steps)[0]
        return steps, S, S_anti

    def _reference_draws_device(self, n, steps):
        key = (self.seed, int(n), int(steps), id(self.handle))
        cache = getattr(self, "_ref_pcg64", None)
        if cache is None or cache[0] != key:
            if cache is not None:
                cache[1].close()
            self._ref_pcg64 = cache = (key, _lib.ReferenceDraws(self.handle, self.seed, n, steps))
        return cache[1]

    def _price_reference(self, spot, strike, T, is_call):
        p, n = self.params, self.num_paths
        steps, S, S_anti = self._terminal_reference(spot, T)
        discount = np.exp(-p.r * T)
        pay = (lambda s: np.maximum(s - strike, 0.0)) if is_call else (lambda s: np.maximum(strike - s, 0.0))
        a = pay(S)
        payoffs = 0.5 * (a + pay(S_anti)) if self.use_antithetic else a
        raw_price = discount * np.mean(payoffs)
        result = {"price": raw_price, "std_error": discount * np.std(payoffs) / np.sqrt(n),
                  "num_paths_used": n, "num_steps": steps}
        if self.use_control_variate:
            bs_ref = bs_price(float(spot), strike, T, p.r, p.q, np.sqrt(p.v0), is_call)
            bs_mc = discount * np.mean(a)
            result["price"] = raw_price - (bs_mc - bs_ref)
            result["bs_cv_adjustment"] = bs_mc - bs_ref
            result["bs_ref"] = bs_ref
            result["raw_mc_price"] = raw_price
            result["std_error"] = discount * np.std(payoffs - (a - bs_ref / discount)) / np.sqrt(n)
        return result

    def _price_batch_reference(self, spot, strikes, T, is_call):
        p, n = self.params, self.num_paths
        _, S, S_anti = self._terminal_reference(spot, T)
        discount = np.exp(-p.r * T)
        sigma_bs = np.sqrt(p.v0)
        results = []
        for K in strikes:
            pay = (lambda s: np.maximum(s - K, 0.0)) if is_call else (lambda s: np.maximum(K - s, 0.0))
            a = pay(S)
            payoffs = 0.5 * (a + pay(S_anti)) if self.use_antithetic else a
            raw = discount * np.mean(payoffs)
            res = {"strike": K, "price": raw, "std_error": discount * np.std(payoffs) / np.sqrt(n)}
            if self.use_control_variate:
                bs_ref = bs_price(float(spot), float(K), T, p.r, p.q, sigma_bs, is_call)
                res["price"] = raw - (discount * np.mean(a) - bs_ref)
                res["bs_ref"] = bs_ref
            results.append(res)
        return results
