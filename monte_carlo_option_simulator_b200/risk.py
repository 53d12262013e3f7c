"""Host mirror of engine/risk.py: the path-based tail metrics (compute_risk_metrics :117-155, Hill :158-173) and the
callers built on MonteCarloEngine.price -- StressTestEngine (:23-111), LiquidityStress (:178-222), HedgingBacktest
(:228-337).

The reference sorts the whole P&L vector on the host; here the vector goes to the GPU once (or already lives there)
and the order statistics come from an exact radix select (csrc/risk.cu).  Same keys, same index conventions.
The reference prices every stress scenario and every hedging scenario's premium with its own price() call; here all
scenarios of a report go out as one (cell x path) launch (MonteCarloEngine.price_many -> b200mc_price_cells) and the
hedging walk runs as one kernel over the scenarios (b200mc_hedge_walk)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np

from . import _lib
from .models import SVJParams, copy_with
from .monte_carlo import MonteCarloEngine

SPOT_SHOCKS = [-0.08, -0.05, -0.02, 0.02, 0.05, 0.08]     # engine/config.py:134
VOL_SHOCKS = [-0.05, 0.05]                                # engine/config.py:135  (+-5 vol points)
JUMP_SCENARIO_SIZE = 0.04                                 # engine/config.py:136

KEYS = ("var", "cvar", "skewness", "kurtosis", "excess_kurtosis", "tail_index", "mean", "std")


def compute_risk_metrics(returns, confidence: float = 0.99, *, handle=None) -> Dict[str, float]:
    h = handle or _lib.default_handle()
    a = np.asarray(returns)
    if a.size == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")     # what the reference raises (:129)
    out = h.risk_metrics(a, confidence)
    return {k: float(v) for k, v in zip(KEYS, out)}


def _key_to_value(key: int) -> float:
    """Inverse of the order-preserving image used by csrc/risk.cu (negative doubles: ~bits, others: bits | 2^63)."""
    bits = (key & 0x7FFFFFFFFFFFFFFF) if key >> 63 else (~key & 0xFFFFFFFFFFFFFFFF)
    return float(np.array([bits], dtype=np.uint64).view(np.float64)[0])


def compute_risk_metrics_sharded(local_returns, confidence: float = 0.99, *, comm, handle=None) -> Dict[str, float]:
    """compute_risk_metrics (engine/risk.py:117-155) over a vector that is SHARDED across the ranks of `comm`
    (each rank passes its own shard: a NumPy array, or (device_ptr, n, dtype)).  The shards never move: every rank runs
    the radix select on its own keys and only 2 doubles, 8 x 512 counters and 6 doubles are all-reduced, so all ranks
    take the same decisions and return the same global metrics (same index conventions as the reference)."""
    h = handle or _lib.default_handle()
    if getattr(comm, "handle", None) is h and hasattr(h, "risk_metrics_sharded"):
        # PeerComm: the whole select runs device-side, histograms all-reduced over NVLink peer memory between the kernels
        if isinstance(local_returns, tuple):
            ptr, n_local, dt = local_returns
            out = h.risk_metrics_sharded(int(ptr), confidence, n=n_local, dtype=dt)
        else:
            out = h.risk_metrics_sharded(np.asarray(local_returns), confidence)
        return {k: float(v) for k, v in zip(KEYS, out)}
    if isinstance(local_returns, tuple):
        ptr, n_local, dt = local_returns
        s = h.risk_begin(int(ptr), n_local, dt)
    else:
        a = np.asarray(local_returns)
        n_local = a.size
        s = h.risk_begin(a)
    tot = comm.allreduce_sum(np.array([s[0], s[1], float(n_local)]))
    n, m = int(round(tot[2])), int(round(tot[1]))
    if n == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    mean = tot[0] / n                                                        # :137
    cutoff = max(int(n * (1 - confidence)), 0)                               # :128
    rank_c = cutoff if cutoff < n else 0                                     # :129
    want_hill = m > 20                                                       # :150
    k = 0
    if want_hill:
        k = min(max(int(math.sqrt(m)), 10), m - 1)                           # :165-166
    nsel = 2 if want_hill else 1
    prefix, ranks = [0, 0], [rank_c, k]
    for radix_pass in range(7, -1, -1):
        hist = comm.allreduce_sum(h.risk_hist(radix_pass, nsel, prefix).astype(np.float64)).reshape(2, 256)
        for sel in range(nsel):
            cum = 0
            chosen = 255
            for b in range(256):
                c = int(round(hist[sel, b]))
                if cum + c > ranks[sel]:
                    chosen = b
                    break
                cum += c
            ranks[sel] -= cum
            prefix[sel] |= chosen << (8 * radix_pass)
    thr = [_key_to_value(prefix[0]), _key_to_value(prefix[1]) if want_hill else 0.0]
    r2 = comm.allreduce_sum(h.risk_finish(mean, nsel, thr))
    sd = math.sqrt(r2[0] / n)                                                # :138
    sdc = max(sd, 1e-10)                                                     # :141
    skew = (r2[1] / n) / sdc ** 3
    kurt = (r2[2] / n) / sdc ** 4
    if cutoff <= 0:
        cvar = -thr[0]
    elif cutoff >= n:
        cvar = -mean
    else:
        cvar = -(r2[4] + (cutoff - r2[3]) * thr[0]) / cutoff                 # ties at the threshold
    tail = float("nan")
    if want_hill and thr[1] < 0.0 and r2[5] > 0.0:
        tail = k / r2[5]                                                     # :168-173
    return {"var": -thr[0], "cvar": float(cvar), "skewness": float(skew), "kurtosis": float(kurt),
            "excess_kurtosis": float(kurt - 3.0), "tail_index": float(tail), "mean": float(mean), "std": float(sd)}


def terminal_pnl_metrics(params, spot: float, strike: float, T: float, n_paths: int, n_steps: int, seed: int = 42,
                         is_call: bool = True, premium: Optional[float] = None, confidence: float = 0.99, *,
                         handle=None, dtype=np.float64) -> Dict[str, float]:
    """BASELINE config 4 without leaving the GPU: terminal spots (fused simulator) -> discounted option P&L
    D * payoff(S_T) - premium -> compute_risk_metrics.  premium defaults to the Monte Carlo price of the same paths."""
    h = handle or _lib.default_handle()
    dt = np.dtype(dtype)
    n = int(n_paths)
    disc = float(np.exp(-params.r * T))
    S = h.malloc(n * dt.itemsize)
    pnl = h.malloc(n * 8)
    try:
        h.simulate_terminal(params, float(spot), float(T), int(n_steps), n, seed, _lib.FP64 if dt == np.float64 else 0, dt,
                            dev_ptrs=(S, None, None))
        if premium is None:
            h.option_pnl(S, n, strike, is_call, disc, 0.0, pnl, dtype_in=dt)
            premium = float(h.risk_metrics(pnl, confidence, n=n, dtype=np.float64)[6])      # mean of D * payoff
        h.option_pnl(S, n, strike, is_call, disc, premium, pnl, dtype_in=dt)
        out = h.risk_metrics(pnl, confidence, n=n, dtype=np.float64)
    finally:
        h.free(S)
        h.free(pnl)
    res = {k: float(v) for k, v in zip(KEYS, out)}
    res["premium"] = float(premium)
    return res


# ------------------------------------------------------------------------------------------------------------
# a11: the callers of engine/risk.py
# ------------------------------------------------------------------------------------------------------------
class StressTestEngine:
    """Mirror of StressTestEngine (engine/risk.py:23-111): same constructor, methods, keys and numbers as the reference
    run on this package's MonteCarloEngine with the reference's default flags; every ladder (and the whole report) is
    ONE launch.  Keyword-only extras: rng, precision, handle, comm (as MonteCarloEngine)."""

    def __init__(self, params, num_paths: int = 200_000, seed: int = 42, *, rng=None, precision="fp32", handle=None,
                 comm=None):
        self.params = params
        self.num_paths = num_paths
        self.seed = seed
        self._kw = dict(rng=rng, precision=precision, handle=handle, comm=comm)

    def _engine(self) -> MonteCarloEngine:
        return MonteCarloEngine(self.params, num_paths=self.num_paths, seed=self.seed, **self._kw)   # :36,56,85

    def _vol_shocked(self, shock: float) -> SVJParams:
        p = self.params                                                                            # :61-68
        return copy_with(p, v0=max(p.v0 + 2 * np.sqrt(p.v0) * shock, 0.001), theta=max(p.theta + shock ** 2, 0.001))

    # every method lists its scenarios, prices them in one launch and formats the reference's dicts
    def _prices(self, spots, params, strike, T, is_call) -> List[float]:
        res = self._engine().price_many(spots, strike, T, is_call, params=params)
        return [r["price"] for r in res]

    @staticmethod
    def _spot_rows(spot, base, prices):
        return [{"shock_pct": sh * 100, "spot": spot * (1 + sh), "price": pr, "pnl": pr - base,
                 "pnl_pct": (pr - base) / max(base, 1e-6) * 100} for sh, pr in zip(SPOT_SHOCKS, prices)]   # :42-49

    def spot_shock_ladder(self, spot: float, strike: float, T: float, is_call: bool = True) -> List[Dict]:
        pr = self._prices([spot] + [spot * (1 + sh) for sh in SPOT_SHOCKS], [self.params] * 7, strike, T, is_call)
        return self._spot_rows(spot, pr[0], pr[1:])

    def _vol_rows(self, base, shocked, prices):
        return [{"vol_shock": sh * 100, "v0": sp.v0, "price": pr, "pnl": pr - base}                 # :71-76
                for sh, sp, pr in zip(VOL_SHOCKS, shocked, prices)]

    def vol_shock_ladder(self, spot: float, strike: float, T: float, is_call: bool = True) -> List[Dict]:
        shocked = [self._vol_shocked(sh) for sh in VOL_SHOCKS]
        pr = self._prices([spot] * (1 + len(shocked)), [self.params] + shocked, strike, T, is_call)
        return self._vol_rows(pr[0], shocked, pr[1:])

    @staticmethod
    def _jump_dict(base, dn, up, gap_size):
        return {"base_price": base, "gap_down_price": dn, "gap_down_pnl": dn - base, "gap_up_price": up,
                "gap_up_pnl": up - base, "gap_size_pct": gap_size * 100}                            # :96-103

    def jump_scenario(self, spot: float, strike: float, T: float, is_call: bool = True,
                      gap_size: float = JUMP_SCENARIO_SIZE) -> Dict:
        pr = self._prices([spot, spot * (1 - gap_size), spot * (1 + gap_size)], [self.params] * 3, strike, T, is_call)
        return self._jump_dict(pr[0], pr[1], pr[2], gap_size)

    def full_stress_report(self, spot: float, strike: float, T: float, is_call: bool = True) -> Dict:
        """All scenarios of the three ladders (:105-111) as one launch; the base price (the reference recomputes it,
        identically, for every ladder) is priced once."""
        shocked = [self._vol_shocked(sh) for sh in VOL_SHOCKS]
        g = JUMP_SCENARIO_SIZE
        spots = [spot] + [spot * (1 + sh) for sh in SPOT_SHOCKS] + [spot] * len(shocked) + [spot * (1 - g), spot * (1 + g)]
        params = [self.params] * 7 + shocked + [self.params] * 2
        pr = self._prices(spots, params, strike, T, is_call)
        ns = len(SPOT_SHOCKS)
        return {"spot_shocks": self._spot_rows(spot, pr[0], pr[1:1 + ns]),
                "vol_shocks": self._vol_rows(pr[0], shocked, pr[1 + ns:1 + ns + len(shocked)]),
                "jump_scenario": self._jump_dict(pr[0], pr[-2], pr[-1], g)}


class LiquidityStress:
    """Mirror of LiquidityStress (engine/risk.py:178-222): closed-form parameter transforms, no simulation."""

    @staticmethod
    def bid_ask_widening(base_spread: float, widening_factor: float = 3.0) -> Dict:
        stressed = base_spread * widening_factor
        return {"base_spread": base_spread, "stressed_spread": stressed, "slippage_increase": stressed - base_spread}

    @staticmethod
    def vol_gap_no_spot_move(params, vol_jump: float = 0.05) -> SVJParams:
        return copy_with(params, v0=params.v0 + 2 * np.sqrt(params.v0) * vol_jump + vol_jump ** 2)      # :201

    @staticmethod
    def expiry_vol_crush(params, crush_pct: float = 0.30) -> SVJParams:
        return copy_with(params, theta=max(params.theta * (1 - crush_pct * 0.5), 0.001),              # :214-221
                            v0=max(params.v0 * (1 - crush_pct), 0.001))


class HedgingBacktest:
    """Mirror of HedgingBacktest (engine/risk.py:228-337).  The premiums of all scenarios are one (cell x path) launch
    (cell s: seed + s, the reference's default engine flags), the daily walk is one kernel over the scenarios.
    rng="reference": premiums from the reference's own draws (Sobol front end on the host) and the walk on
    default_rng(seed) normals -- reproduces the reference's numbers; rng="philox" (default): device draws."""

    def __init__(self, params, seed: int = 42, *, rng=None, precision="fp32", handle=None, comm=None):
        self.params = params
        self.seed = seed
        self.rng = rng
        self.comm = comm            # scenarios are sharded over its ranks (contiguous ranges), P&Ls gathered once
        self._kw = dict(rng=rng, precision=precision, handle=handle)

    def run_backtest(self, spot: float, strike: float, T: float, is_call: bool = True, num_days: int = None,
                     txn_cost_bps: float = 5.0, slippage_bps: float = 2.0, num_scenarios: int = 1000,
                     num_mc_paths: int = 50_000) -> Dict:
        if num_days is None:
            num_days = max(int(T * 252), 1)                                                        # :255-256
        p = self.params
        engine = MonteCarloEngine(p, num_paths=num_mc_paths, seed=self.seed, **self._kw)
        h = engine.handle
        lo, hi = 0, num_scenarios
        world = self.comm.world if self.comm is not None else 1
        if world > 1:
            from .dist import shard_range
            lo, hi = shard_range(num_scenarios, self.comm.rank, world)
        pnl = np.zeros(num_scenarios)
        cost = np.zeros(num_scenarios)
        if hi > lo:
            seeds = [self.seed + s for s in range(lo, hi)]                                         # :271
            premiums = engine.prices_for_seeds(spot, strike, T, is_call, seeds)                    # :272-273
            Z = None
            if engine.rng == "reference":                                                          # :262,292
                Z = np.random.default_rng(self.seed).standard_normal((num_scenarios, num_days))[lo:hi]
            pnl[lo:hi], cost[lo:hi] = h.hedge_walk(p, spot, strike, T, is_call, num_days, hi - lo, txn_cost_bps + slippage_bps,
                                                   premiums, Z, seed=self.seed, scenario_offset=lo)
        if world > 1:
            both = self.comm.allreduce_sum(np.stack([pnl, cost]))
            pnl, cost = both.reshape(2, num_scenarios)
        metrics = compute_risk_metrics(pnl, confidence=0.99, handle=h)                             # :319
        return {"mean_pnl": float(np.mean(pnl)), "std_pnl": float(np.std(pnl)),
                "pnl_percentiles": {f"{q}%": float(np.percentile(pnl, q)) for q in (1, 5, 25, 50, 75, 95, 99)},
                "risk_metrics": metrics, "num_scenarios": num_scenarios,
                "total_txn_cost_avg": float(cost[-1])}      # :336: the reference reports the LAST scenario's total
