"""Golden implied-volatility surface written by the reference's own extract_iv_surface / implied_vol
(engine/surface.py:48-126) on a synthetic option chain:

    python tests/golden/make_iv_golden.py      ->  tests/golden/iv_golden.npz

The chain: 6 maturities x 17 strikes around spot 22500, call and put mids from Black-Scholes with a skewed smile, plus
cells built to exercise every branch -- prices below intrinsic and above the hi = 5.0 bound (no root -> None), zero
prices, one NaN, wide bid-ask spreads (filtered), a near-zero maturity.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from engine.surface import bs_call_price, bs_put_price, extract_iv_surface, implied_vol   # noqa: E402  (the reference itself)

spot, r, q = 22500.0, 0.065, 0.012
strikes = np.linspace(0.7, 1.3, 17) * spot
mats = np.array([1e-11, 0.02, 0.08, 0.25, 0.5, 1.0])
g = np.random.default_rng(5)
calls = np.zeros((mats.size, strikes.size))
puts = np.zeros_like(calls)
true_iv = np.zeros_like(calls)
for i, T in enumerate(mats):
    for j, K in enumerate(strikes):
        m = np.log(K / spot)
        iv = 0.16 + 0.25 * m * m - 0.12 * m + 0.03 * np.sqrt(max(T, 1e-3))
        true_iv[i, j] = iv
        calls[i, j] = bs_call_price(spot, K, T, r, q, iv)
        puts[i, j] = bs_put_price(spot, K, T, r, q, iv)
# branches
calls[2, 0] = 0.5 * max(spot * np.exp(-q * mats[2]) - strikes[0] * np.exp(-r * mats[2]), 0.0)   # below intrinsic: no root
puts[3, 16] *= 0.2                                                                               # below intrinsic
calls[4, 8] = 1.2 * spot                                                                         # above any BS price
puts[1, 5] = 0.0
calls[5, 3] = np.nan
spreads = 0.01 * 0.5 * (calls + puts)
spreads[3, 4] = 0.5 * 0.5 * (calls[3, 4] + puts[3, 4])                                           # 50 % spread: filtered
spreads[4, 12] = 0.2 * 0.5 * (calls[4, 12] + puts[4, 12])
surf = extract_iv_surface(spot, r, q, strikes, mats, calls, puts, spreads)
surf_nospread = extract_iv_surface(spot, r, q, strikes, mats, calls, puts)
# scalar calls with non-default bounds
scal = []
for price, K, T, call, lo, hi in [(900.0, 22500.0, 0.25, True, 0.001, 5.0), (900.0, 22500.0, 0.25, False, 0.001, 5.0),
                                  (50.0, 25000.0, 0.1, True, 0.05, 2.0), (5000.0, 22500.0, 0.25, True, 0.001, 0.5),
                                  (1e-6, 30000.0, 0.05, True, 0.001, 5.0)]:
    v = implied_vol(price, spot, K, T, r, q, call, lo, hi)
    scal.append([price, K, T, float(call), lo, hi, np.nan if v is None else v])
np.savez_compressed(os.path.join(HERE, "iv_golden.npz"), spot=spot, r=r, q=q, strikes=strikes, maturities=mats, calls=calls,
                    puts=puts, spreads=spreads, true_iv=true_iv, iv_call=surf["iv_call"], iv_put=surf["iv_put"],
                    valid=surf["valid_mask"], iv_call_ns=surf_nospread["iv_call"], iv_put_ns=surf_nospread["iv_put"],
                    valid_ns=surf_nospread["valid_mask"], scalar=np.array(scal))
print("wrote iv_golden.npz: valid", int(surf["valid_mask"].sum()), "of", surf["valid_mask"].size,
      "; NaN calls", int(np.isnan(surf["iv_call"]).sum()), "puts", int(np.isnan(surf["iv_put"]).sum()))
print(np.array(scal)[:, -1])
