"""Randomised cross-check of b200mc_price_cells against single-problem launches: many launches with random numbers of
cells, strikes, modes, path counts (1 .. 400k, so 1 .. 1500 batches per cell), steps, offsets, flags.
    python tools/fuzz_cells.py [n_launches] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 150
g = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
h = _lib.Handle(0)
worst = 0.0
for it in range(n_launch):
    n_cells = int(g.choice([1, 2, 5, 17, 64, 300]))
    ks_n = int(g.choice([1, 1, 2, 5, 21, 64]))
    fl = (int(g.integers(0, 2)) * _lib.ANTITHETIC) | (int(g.integers(0, 2)) * _lib.FP64)
    big = g.random() < 0.2
    cells, strikes = [], []
    for c in range(n_cells):
        mode = g.choice(["gbm", "detvar", "heston", "svj"])
        v0 = float(g.uniform(0.01, 0.3))
        kw = dict(kappa=0.0, theta=v0, xi=0.0, rho=float(g.uniform(-0.9, 0.9)), v0=v0, lambda_j=0.0, mu_j=-0.05, sigma_j=0.1,
                  r=float(g.uniform(0, 0.1)), q=float(g.uniform(0, 0.04)))
        if mode != "gbm":
            kw.update(kappa=float(g.uniform(0.3, 5)), theta=float(g.uniform(0.01, 0.3)))
        if mode in ("heston", "svj"):
            kw.update(xi=float(g.uniform(0.1, 1.0)))
        if mode == "svj":
            kw.update(lambda_j=float(g.uniform(0.3, 6)))
        S0 = float(g.uniform(10, 30000))
        npaths = int(g.integers(1, 400_000 if (big and c == 0) else 3000))
        cells.append(dict(params=SVJParams(**kw), S0=S0, T=float(g.uniform(0.05, 2)), n_steps=int(g.integers(1, 80)),
                          n_paths=npaths, seed=int(g.integers(0, 2 ** 63)), path_offset=int(g.choice([0, 3, 2 ** 32 - 50, 2 ** 45])),
                          is_call=bool(g.integers(0, 2))))
        strikes.append(np.sort(S0 * g.uniform(0.6, 1.4, ks_n)))
    got = h.price_cells(cells, np.array(strikes), fl)
    for idx in set([0, n_cells - 1] + list(g.integers(0, n_cells, size=min(n_cells, 6)))):
        c = cells[idx]
        want = h.price_european(c["params"], c["S0"], c["T"], c["n_steps"], c["n_paths"], c["seed"], strikes[idx], c["is_call"], fl,
                                None, path_offset=c["path_offset"])
        scale = np.maximum(np.abs(want[:, :9]), 1e-9 * max(1.0, c["S0"]) ** 2)
        err = float(np.max(np.abs(got[idx][:, :9] - want[:, :9]) / scale))
        worst = max(worst, err)
        assert err < 1e-10, (it, idx, err, c, fl, ks_n)
        assert not got[idx][:, 9:].any()
print(f"{n_launch} launches ok, worst relative deviation from the single-problem launches {worst:.2e}")
h.close()
