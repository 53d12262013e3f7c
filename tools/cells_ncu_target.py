"""One (cell x path) launch of the hedging-backtest shape (1000 cells x 50k paths x 63 steps, GBM, antithetic) and one of
the QMC front end (64k paths x 250 steps): the targets of the ncu captures in profiles/."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

h = _lib.Handle(0)
p = SVJParams.gbm(0.3, r=0.065)
cells = _lib.make_cells(p, 22500.0, 0.25, 63, 50_000, 42 + np.arange(1000))
for _ in range(2):
    out = h.price_cells(cells, np.full(1000, 22500.0), _lib.ANTITHETIC)
print("cells ok", out[0, 0, :3])
t = _lib.sobol_tables(250, 42)
for _ in range(2):
    rows = h.price_european_qmc(p, 2500.0, 1.0, 250, 65536, t, [2500.0])
print("qmc ok", rows[0, :3])
h.close()
