"""Host mirror of the implied-volatility part of engine/surface.py (bs_call_price :22-28, bs_put_price :31-37, bs_vega
:40-45, implied_vol :48-66, extract_iv_surface :69-126) -- SURVEY.md 8(f)-3.

The reference inverts every option with SciPy's brentq on a Python objective (two scipy.stats.norm.cdf calls per
evaluation, ~1 ms per option); here a whole chain -- the (n_mat, n_k) call and put grids extract_iv_surface takes, or the
21 strikes of /api/smile -- is ONE launch of b200mc_implied_vol.  Same signatures, same keys, None / NaN / valid_mask
conventions of the reference.  SABR, spline fitting and the arbitrage checks of surface.py are not part of the hot path
and are not mirrored."""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np

from . import _lib


def _ncdf(x: float) -> float:
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def bs_call_price(S, K, T, r, q, sigma):
    """engine/surface.py:22-28."""
    if T <= 1e-10 or sigma <= 1e-10:
        return max(S * math.exp(-q * T) - K * math.exp(-r * T), 0.0)
    d1 = (math.log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / (sigma * math.sqrt(T))
    d2 = d1 - sigma * math.sqrt(T)
    return S * math.exp(-q * T) * _ncdf(d1) - K * math.exp(-r * T) * _ncdf(d2)


def bs_put_price(S, K, T, r, q, sigma):
    """engine/surface.py:31-37."""
    if T <= 1e-10 or sigma <= 1e-10:
        return max(K * math.exp(-r * T) - S * math.exp(-q * T), 0.0)
    d1 = (math.log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / (sigma * math.sqrt(T))
    d2 = d1 - sigma * math.sqrt(T)
    return K * math.exp(-r * T) * _ncdf(-d2) - S * math.exp(-q * T) * _ncdf(-d1)


def bs_vega(S, K, T, r, q, sigma):
    """engine/surface.py:40-45."""
    if T <= 1e-10 or sigma <= 1e-10:
        return 0.0
    d1 = (math.log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / (sigma * math.sqrt(T))
    return S * math.exp(-q * T) * math.sqrt(T) * math.exp(-0.5 * d1 * d1) / math.sqrt(2.0 * math.pi)


def implied_vol_batch(prices, S: float, strikes, maturities, r: float, q: float, is_call=True, lo: float = 0.001,
                      hi: float = 5.0, *, handle=None) -> np.ndarray:
    """NEW: implied_vol for arrays (broadcast); NaN where the scalar function returns None."""
    return (handle or _lib.default_handle()).implied_vol(prices, S, strikes, maturities, r, q, is_call, lo, hi)


def implied_vol(price: float, S: float, K: float, T: float, r: float, q: float, is_call: bool = True,
                lo: float = 0.001, hi: float = 5.0, *, handle=None) -> Optional[float]:
    """engine/surface.py:48-66: the root of BS(sigma) - price in [lo, hi], None when there is none."""
    try:
        iv = float(implied_vol_batch([float(price)], float(S), [float(K)], [float(T)], float(r), float(q), bool(is_call),
                                     lo, hi, handle=handle)[0])
    except (TypeError, ValueError):
        return None
    return None if math.isnan(iv) else iv


def extract_iv_surface(spot: float, r: float, q: float, strikes: np.ndarray, maturities: np.ndarray,
                       call_prices: np.ndarray, put_prices: np.ndarray, bid_ask_spreads: Optional[np.ndarray] = None,
                       max_spread_pct: float = 0.10, *, handle=None) -> Dict:
    """engine/surface.py:69-126, all 2 n_mat n_k inversions in one launch."""
    strikes = np.asarray(strikes)
    maturities = np.asarray(maturities)
    call_prices = np.asarray(call_prices, dtype=np.float64)
    put_prices = np.asarray(put_prices, dtype=np.float64)
    n_mat, n_k = call_prices.shape
    valid = np.ones((n_mat, n_k), dtype=bool)
    liquid = np.ones((n_mat, n_k), dtype=bool)
    if bid_ask_spreads is not None:                                                     # :98-103
        mid = 0.5 * (call_prices + put_prices)
        with np.errstate(divide="ignore", invalid="ignore"):
            liquid = ~((mid > 0) & (np.asarray(bid_ask_spreads, dtype=np.float64) / mid > max_spread_pct))
    both = np.stack([call_prices, put_prices])                                          # [2, n_mat, n_k]
    flags = np.array([True, False])[:, None, None]
    iv = implied_vol_batch(both, spot, np.asarray(strikes, dtype=np.float64)[None, None, :],
                           np.asarray(maturities, dtype=np.float64)[None, :, None], r, q, flags, handle=handle)
    iv[:, ~liquid] = np.nan                                                             # skipped cells stay NaN (:102-103)
    valid &= liquid & ~np.isnan(iv[0]) & ~np.isnan(iv[1])                               # :108-116
    return {"iv_call": iv[0], "iv_put": iv[1], "valid_mask": valid, "strikes": strikes, "maturities": maturities}


def iv_surface_from_engine(engine, spot: float, strikes, maturities) -> Dict:
    """NEW: Monte Carlo call and put grids (MonteCarloEngine.price_grid, the (n_mat, n_k) layout above) -> IV surface,
    everything on the engine's device."""
    calls = engine.price_grid(spot, strikes, maturities, is_call=True)
    puts = engine.price_grid(spot, strikes, maturities, is_call=False)
    out = extract_iv_surface(spot, engine.params.r, engine.params.q, np.asarray(strikes, dtype=np.float64),
                             np.asarray(maturities, dtype=np.float64), calls["prices"], puts["prices"], handle=engine.handle)
    out.update(call_prices=calls["prices"], put_prices=puts["prices"], num_steps=calls["num_steps"])
    return out
