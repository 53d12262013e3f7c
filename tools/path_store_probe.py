"""Path-store kernel: GB/s for several row pitches / dtypes (one line each).  python tools/path_store_probe.py [n_paths]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
h = _lib.Handle(0)
p = SVJParams.gbm(0.3, r=0.065)
for name, dt, fl, esz, ld in (("f32 ld251", np.float32, 0, 4, 251), ("f32 ld252", np.float32, 0, 4, 252),
                              ("f32 ld256", np.float32, 0, 4, 256), ("f64 ld251", np.float64, _lib.FP64, 8, 251),
                              ("f64 ld252", np.float64, _lib.FP64, 8, 252), ("f32state f64out", np.float64, 0, 8, 251)):
    buf = torch.empty(n * ld * esz, dtype=torch.uint8, device="cuda")
    best = 1e9
    for r in range(reps + 1):
        h.timer_begin()
        h.generate_paths(p, 2500.0, 1.0, 250, n, 42 + r, fl, dt, 0, ld, out_dev=buf.data_ptr())
        ms = h.timer_end()
        if r:
            best = min(best, ms)
    print(f"{name:16s} {n} paths: {best:8.3f} ms  {n * 251 * esz / best / 1e6:8.1f} GB/s  {n * 250 / best / 1e6:8.1f} Gpath-steps/s", flush=True)
    del buf
h.close()
