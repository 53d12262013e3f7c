"""Tail-metric kernel timing on a device-resident P&L vector (BASELINE cfg4: 4e6 terminal P&Ls)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

h = _lib.Handle(0)
for n in (4_000_000, 40_000_000):
    x = torch.from_numpy(np.random.default_rng(0).standard_t(4, size=n) * 0.01).cuda()
    h.risk_metrics(x.data_ptr(), 0.99, n=n, dtype=np.float64)
    t0 = time.perf_counter()
    for _ in range(10):
        out = h.risk_metrics(x.data_ptr(), 0.99, n=n, dtype=np.float64)
    dt = (time.perf_counter() - t0) / 10
    xs = x.cpu().numpy()
    t0 = time.perf_counter()
    srt = np.sort(xs)
    tn = time.perf_counter() - t0
    print(f"n={n}: device risk metrics {dt * 1e3:.3f} ms/call ({n * 8 * 10 / dt / 1e9:.0f} GB/s over 10 passes of n*8 B); np.sort alone {tn * 1e3:.0f} ms; var={out[0]:.6f}")
h.close()
