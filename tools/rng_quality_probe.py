"""Joint law of consecutive normals on a 64 x 64 grid of equiprobable cells: production layout (one Box-Muller pair per
32-bit word: radius and angle share 14 bits) vs the validation twin (B200MC_WIDE_RNG: a pair per two words), same-word
pairs (lag 0) and adjacent-word pairs (lag 1), at growing sample sizes.  Prints the z-score of the chi-square, the rms
relative deviation of the cell probabilities beyond sampling noise, marginal chi-squares and Kolmogorov distances; then
prices the same options with both layouts."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib, bs_price  # noqa: E402

h = _lib.Handle(0)


def stats(c):
    n = c.sum()
    e = n / c.size
    chi2 = ((c - e) ** 2).sum() / e
    z = (chi2 - (c.size - 1)) / math.sqrt(2 * (c.size - 1))
    excess = max(chi2 / (c.size - 1) - 1.0, 0.0) * (c.size / n)          # variance of the cell probabilities beyond Poisson noise
    out = [z, math.sqrt(excess)]
    for m in (c.sum(axis=1), c.sum(axis=0)):
        em = n / 64
        out.append((((m - em) ** 2).sum() / em - 63) / math.sqrt(126))
        out.append(np.abs(np.cumsum(m) / n - np.arange(1, 65) / 64.0).max() * math.sqrt(n))
    return out


print("layout         lag   pairs       chi2 z    rms cell dev   marg1 z  KS1*sqrtN  marg2 z  KS2*sqrtN")
for name, wide in (("production", 0), ("round-1 layout", _lib.HIST_R01), ("wide twin", _lib.HIST_WIDE)):
    for lag in (0, 1):
        for paths, blocks in ((250_000, 10), (2_500_000, 10), (4_000_000, 32)):
            c = h.normal_hist2d(42, paths, blocks, lag | wide).astype(np.float64)
            s = stats(c)
            print(f"{name:14s} {lag:3d} {int(c.sum()):11d} {s[0]:10.2f} {s[1]:12.3e} {s[2]:9.2f} {s[3]:9.3f} {s[4]:9.2f} {s[5]:9.3f}", flush=True)

p = SVJParams.gbm(0.30, r=0.065, q=0.0)
for steps, T, n in ((10, 0.04, 1_000_000_000), (250, 1.0, 200_000_000)):
    sd = 0.30 * math.sqrt(T)
    for mny in (0.0, 2.0, 3.5):
        K = 2500.0 * math.exp(mny * sd)
        res = []
        for fl in (0, _lib.WIDE_RNG):
            row = h.price_european(p, 2500.0, T, steps, n, 2024, [K], True, fl, None)[0]
            mean, m2 = row[1] / n, row[3] / n
            disc = math.exp(-p.r * T)
            res.append((disc * mean, disc * math.sqrt(max(m2 - mean * mean, 0) / n)))
        bs = bs_price(2500.0, K, T, p.r, p.q, 0.30, True)
        zz = (res[0][0] - res[1][0]) / math.hypot(res[0][1], res[1][1])
        print(f"{steps:4d} steps, {n:.0e} paths, strike +{mny} sd: production {res[0][0]:.6f} +- {res[0][1]:.6f}, wide twin {res[1][0]:.6f} +- "
              f"{res[1][1]:.6f}, Black-Scholes {bs:.6f}; z(prod - BS) {(res[0][0] - bs) / res[0][1]:+.2f}, z(wide - BS) "
              f"{(res[1][0] - bs) / res[1][1]:+.2f}, z(prod - wide) {zz:+.2f}", flush=True)
h.close()
