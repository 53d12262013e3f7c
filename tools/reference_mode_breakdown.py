"""Where a rng="reference", use_sobol=False call spends its time (50k x 250): allocation, NumPy draws on the device, the
recurrence kernel, read-backs, host reductions."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import GreeksEngine, MonteCarloEngine, SVJParams, _lib  # noqa: E402

h = _lib.Handle(0)
p = SVJParams()
n, steps = 50_000, 250
N = n * steps


def t(label, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        h.synchronize()
        t0 = time.perf_counter()
        r = fn()
        h.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{label:60s} {best * 1e3:8.3f} ms", flush=True)
    return r


buf = t("h.malloc(4 * N * 8 + ...)", lambda: h.malloc(4 * N * 8 + 16 * n * 8), 1)
t("h.free", lambda: h.free(buf), 1)
d = t("ReferenceDraws(seed) [malloc + normals + uniforms]", lambda: _lib.ReferenceDraws(h, 42, n, steps), 3)
t("d.simulate (kernel + 2 read-backs)", lambda: d.simulate(p, 22500.0, 1.0))
t("d.simulate negate", lambda: d.simulate(p, 22500.0, 1.0, negate=True))
t("h.simulate_given_normals_dev only", lambda: h.simulate_given_normals_dev(p, 22500.0, 1.0, n, steps, d.Z1, d.Z2, d.Zj, d.Zjs, d.S, d.v))
S = np.empty(n)
t("h.d2h(S) 400 KB", lambda: h.d2h(S, d.S))
t("numpy payoff + mean + std on 50k", lambda: (np.maximum(S - 22500.0, 0).mean(), np.std(S)))
e = MonteCarloEngine(p, n, 250, 42, use_sobol=False, use_antithetic=True, use_control_variate=True, rng="reference", handle=h)
t("price() first (draws generated)", lambda: e.price(22500.0, 22500.0, 1.0), 1)
t("price() again (draws cached)", lambda: e.price(22500.0, 22500.0, 1.0))
g = GreeksEngine(p, n, 250, 42, rng="reference", handle=h)
t("delta() first", lambda: g.delta(22500.0, 22500.0, 1.0), 1)
t("delta() again", lambda: g.delta(22500.0, 22500.0, 1.0))
t("vega()", lambda: g.vega(22500.0, 22500.0, 1.0))
t("gamma()", lambda: g.gamma(22500.0, 22500.0, 1.0))
def fresh():
    e2 = MonteCarloEngine(p, n, 250, int(time.time() * 1e6) % 100000, use_sobol=False, use_antithetic=True, use_control_variate=True,
                          rng="reference", handle=h)
    return e2.price(22500.0, 22500.0, 1.0)
t("fresh engine + seed: price()", fresh, 3)
h.close()
