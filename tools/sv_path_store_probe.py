import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monte_carlo_option_simulator_b200 import SVJParams, _lib
h = _lib.Handle(0)
n = 1_000_000
for name, p in (("heston", SVJParams(lambda_j=0.0)), ("svj", SVJParams())):
    for dname, dt, fl, esz in (("f32", np.float32, 0, 4), ("f64out/f32state", np.float64, 0, 8)):
        buf = torch.empty(n * 251 * esz, dtype=torch.uint8, device="cuda")
        best = 1e9
        for r in range(3):
            h.timer_begin(); h.generate_paths(p, 22500.0, 1.0, 250, n, 42 + r, fl, dt, 0, 251, out_dev=buf.data_ptr()); ms = h.timer_end()
            if r: best = min(best, ms)
        print(f"{name:7s} {dname:16s} {n} x 251: {best:.3f} ms  {n*251*esz/best/1e6:.0f} GB/s  {n*250/best/1e6:.0f} Gpath-steps/s")
        del buf
