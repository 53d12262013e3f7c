"""Every kernel family once at a small size, for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

h = _lib.Handle(0)
g = SVJParams.gbm(0.3, r=0.065)
det = SVJParams(kappa=3.0, theta=0.05, xi=0.0, rho=0.0, v0=0.09, lambda_j=0.0, mu_j=0.0, sigma_j=0.0)
hes = SVJParams(lambda_j=0.0)
svj = SVJParams()
bumps = _lib.Bumps(0.01, 0.05, 0.03, 0.0651, 0.0649)
rng = np.random.default_rng(0)
for p in (g, det, hes, svj):
    for fl in (0, _lib.ANTITHETIC, _lib.FP64 | _lib.ANTITHETIC):
        h.price_european(p, 100.0, 0.5, 37, 1500, 1, [100.0], True, fl)
        h.price_european(p, 100.0, 0.5, 37, 1500, 1, [90.0, 100.0, 110.0], True, fl)
        h.simulate_terminal(p, 100.0, 0.5, 37, 1500, 1, fl, np.float32, 0, bool(fl & _lib.ANTITHETIC), True)
    h.price_european(p, 100.0, 0.5, 37, 1500, 1, [100.0], True, _lib.GREEKS, bumps)
    h.price_european(p, 100.0, 0.5, 37, 1500, 1, [95.0, 105.0], True, _lib.GREEKS | _lib.ANTITHETIC, bumps)
    for dt, fl in ((np.float32, 0), (np.float64, _lib.FP64)):
        h.generate_paths(p, 100.0, 0.5, 37, 333, 2, fl, dt)
        h.generate_paths(p, 100.0, 0.5, 37, 333, 2, fl, dt, ld=40)
    for w in range(4):
        h.dump_normals(3, 100, 37, _lib.STREAM_SVJ, w, jump_prob=0.01)
Z = [rng.standard_normal((257, 50)) for _ in range(3)]
h.simulate_given_normals(svj, 100.0, 0.5, Z[0], Z[1], rng.random((257, 50)), Z[2], 50, True)
h.risk_metrics(rng.standard_t(3, size=100_000) * 0.01, 0.99)
h.risk_metrics((rng.standard_t(3, size=5_000) * 0.01).astype(np.float32), 0.95)
h.dump_philox(1, 100, 5, 0)
h.normal_moments(1, 1000, 4)
print("sanitize smoke done, launches:", h.launches)
h.close()
