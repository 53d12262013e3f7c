"""fp32-state path store (BASELINE cfg4, 4M x 251): the MUFU form (ex2 per stored value) against the polynomial /
multiplicative form (DEG 4 / 5) at 3-4 resident CTAs per SM -- GB/s and accuracy against the fp64-state matrix of the same
draws.  The knobs are the environment variables read by b200mc_generate_paths (csrc/paths.cu)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
h = _lib.Handle(0)
p = SVJParams.gbm(0.3, r=0.065)
ref = h.generate_paths(p, 2500.0, 1.0, 250, 20_000, 42, _lib.FP64, np.float64)          # fp64 state, same draws
for expm in ("mufu", "poly4", "poly5"):
    for minb in ("1", "4"):
        os.environ["B200MC_PATHS_EXP"], os.environ["B200MC_PATHS_MINB"] = expm, minb
        got = h.generate_paths(p, 2500.0, 1.0, 250, 20_000, 42, 0, np.float32)
        err = np.abs(got.astype(np.float64) / ref - 1.0)
        line = f"exp={expm:5s} minb={minb}: max rel err {err.max():.2e} (99.9% {np.quantile(err, 0.999):.2e})"
        for name, dt, esz in (("f32", np.float32, 4), ("f64out", np.float64, 8)):
            buf = torch.empty(n * 251 * esz, dtype=torch.uint8, device="cuda")
            best = 1e9
            for r in range(5):
                h.timer_begin()
                h.generate_paths(p, 2500.0, 1.0, 250, n, 42 + r, 0, dt, 0, 251, out_dev=buf.data_ptr())
                ms = h.timer_end()
                if r:
                    best = min(best, ms)
            line += f" | {name} {best:7.3f} ms {n * 251 * esz / best / 1e6:7.1f} GB/s {n * 250 / best / 1e9:6.3f}e12 steps/s"
            del buf
        print(line, flush=True)
h.close()
