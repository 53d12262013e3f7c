"""One launch (after one warm-up launch) of every kernel this round's profiles/ summaries cover; run under
`ncu --set full -k regex:<name> -c <count>` (see tools/gpu_call1.sh).  python tools/ncu_targets.py [target ...]
Targets: gbm32 gbm64 heston svj paths32 paths64 given risk risk40m numpy hedge qmc."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

want = set(sys.argv[1:]) or {"gbm32", "gbm64", "heston", "svj", "paths32", "paths64", "given", "risk", "hedge", "qmc"}
h = _lib.Handle(0)
g = SVJParams.gbm(0.3, r=0.065)
bumps = _lib.Bumps(0.01, g.v0 + 0.01, g.v0 - 0.01, g.r + 1e-4, g.r - 1e-4)
out = h.malloc(17 * 8 * 256)
REPS = int(os.environ.get("NCU_TARGET_REPS", "2"))        # 1 under ncu: every captured launch is replayed anyway


def fused(p, s0, n, fl, b=None):
    for r in range(REPS):
        h.price_european(p, s0, 1.0, 250, n, 42 + r, [s0], True, fl, b, out_dev=out)
    h.synchronize()


if "gbm32" in want:
    fused(g, 2500.0, 10_000_000, _lib.GREEKS, bumps)
if "gbm64" in want:
    fused(g, 2500.0, 10_000_000, _lib.GREEKS | _lib.FP64, bumps)
if "heston" in want:
    fused(SVJParams(lambda_j=0.0), 22500.0, 2_500_000, _lib.ANTITHETIC)
if "svj" in want:
    fused(SVJParams(), 22500.0, 2_500_000, _lib.ANTITHETIC)
for key, dt, esz in (("paths32", np.float32, 4), ("paths64", np.float64, 8)):
    if key in want:
        n = 4_000_000
        buf = torch.empty(n * 251 * esz, dtype=torch.uint8, device="cuda")
        for r in range(REPS):
            h.generate_paths(g, 2500.0, 1.0, 250, n, 42 + r, 0, dt, 0, 251, out_dev=buf.data_ptr())
        h.synchronize()
        del buf
if "given" in want:
    n, steps = 400_000, 250
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    d = [torch.randn(n, steps, dtype=torch.float64, device="cuda", generator=gen) for _ in range(2)]
    d.append(torch.rand(n, steps, dtype=torch.float64, device="cuda", generator=gen))
    d.append(torch.randn(n, steps, dtype=torch.float64, device="cuda", generator=gen))
    S = torch.empty(n, dtype=torch.float64, device="cuda")
    V = torch.empty_like(S)
    sp = _lib.to_params(SVJParams())
    for r in range(REPS):
        h._check(h.lib.b200mc_simulate_given_normals_dev(h.h, C.byref(sp), 2500.0, 1.0, n, steps, *(C.c_void_p(t.data_ptr()) for t in d),
                                                        0, C.c_void_p(S.data_ptr()), C.c_void_p(V.data_ptr()), None))
    h.synchronize()
    del d
if "risk" in want or "risk40m" in want:
    n = 40_000_000 if "risk40m" in want else 4_000_000
    x = torch.from_numpy(np.random.default_rng(0).standard_t(4, size=n) * 0.01).cuda()
    for r in range(REPS):
        h.risk_metrics(x.data_ptr(), 0.99, n=n, dtype=np.float64)
if "numpy" in want:
    buf = h.malloc(3 * 50_000 * 250 * 8)
    for r in range(REPS):
        h.numpy_fill(42, 3 * 50_000 * 250, out_dev=buf)
    h.free(buf)
if "hedge" in want:
    from monte_carlo_option_simulator_b200.risk import HedgingBacktest
    bt = HedgingBacktest(SVJParams(), seed=42, handle=h)
    for r in range(REPS):
        bt.run_backtest(22500.0, 22500.0, 0.25)
if "qmc" in want:
    t = _lib.sobol_tables(250, 42)
    for r in range(REPS):
        h.price_european_qmc(g, 2500.0, 1.0, 250, 65536, t, [2500.0])
print("targets done:", sorted(want), "launches", h.launches)
h.close()
