"""Deterministic-mode kernel (drop-in for _simulate_svj_paths_numba) end to end from HOST arrays at BASELINE cfg1 size,
next to the oracle's OpenMP C port of the same recurrence on this box's host cores."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402

h = _lib.Handle(0)
n, steps = 50_000, 250
g = np.random.default_rng(0)
Z1, Z2, Zjs = (g.standard_normal((n, steps)) for _ in range(3))
Zj = g.random((n, steps))
for name, p in (("gbm (only Z1 is read)", SVJParams.gbm(0.3)), ("heston (Z1, Z2)", SVJParams(lambda_j=0.0)), ("svj (all four)", SVJParams())):
    h.simulate_given_normals(p, 2500.0, 1.0, Z1, Z2, Zj, Zjs, steps)
    t0 = time.perf_counter()
    for _ in range(5):
        S, v, _ = h.simulate_given_normals(p, 2500.0, 1.0, Z1, Z2, Zj, Zjs, steps)
    dt = (time.perf_counter() - t0) / 5
    op = O.as_params(p)
    O._sim(op, 2500.0, 1.0, Z1, Z2, Zj, Zjs, steps)
    t0 = time.perf_counter()
    for _ in range(3):
        So = O._sim(op, 2500.0, 1.0, Z1, Z2, Zj, Zjs, steps)[0]
    dto = (time.perf_counter() - t0) / 3
    print(f"{name:24s} GPU end to end from host arrays {dt * 1e3:7.1f} ms ({n * steps / dt:.2e} path-steps/s) | "
          f"oracle C port, {O.num_threads()} threads {dto * 1e3:7.1f} ms ({n * steps / dto:.2e}) | max rel diff {np.max(np.abs(S / So - 1)):.1e}")
# device-resident inputs: the kernel alone
import torch
d = [torch.from_numpy(a).cuda() for a in (Z1, Z2, Zj, Zjs)]
S = torch.empty(n, dtype=torch.float64, device="cuda"); V = torch.empty_like(S)
import ctypes as C
sp = _lib.to_params(SVJParams())
for _ in range(2):
    h.timer_begin()
    h._check(h.lib.b200mc_simulate_given_normals_dev(h.h, C.byref(sp), 2500.0, 1.0, n, steps, *(C.c_void_p(t.data_ptr()) for t in d), 0,
                                                    C.c_void_p(S.data_ptr()), C.c_void_p(V.data_ptr()), None))
    ms = h.timer_end()
print(f"svj, inputs resident in HBM: kernel {ms:.3f} ms = {4 * n * steps * 8 / ms / 1e6:.0f} GB/s of input read")
