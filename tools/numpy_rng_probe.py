"""NumPy's PCG64 Ziggurat normals / uniforms generated on the device (csrc/np_normal.cu) at the reference's size:
time for the whole front end (3 x 50k x 250 normals + 50k x 250 uniforms, device resident), bitwise check against NumPy,
and rng="reference" use_sobol=False price() / GreeksEngine end to end vs the host front end."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import GreeksEngine, MonteCarloEngine, SVJParams, _lib  # noqa: E402

h = _lib.Handle(0)
n, steps = 50_000, 250
N = n * steps
buf = h.malloc(4 * N * 8)
for seed in (42, 7):
    t0 = time.perf_counter()
    g = np.random.default_rng(seed)
    want = [g.standard_normal((n, steps)) for _ in range(3)]
    wu = np.random.default_rng(seed + 1).random((n, steps))
    t_np = time.perf_counter() - t0
    h.numpy_fill(seed, 3 * N, out_dev=buf)
    h.synchronize()
    best = 1e9
    for _ in range(5):
        l0 = h.launches
        t0 = time.perf_counter()
        _, used = h.numpy_fill(seed, 3 * N, out_dev=buf)
        h.numpy_fill(seed + 1, N, _lib.NUMPY_RANDOM, out_dev=buf + 3 * N * 8)
        h.synchronize()
        best = min(best, time.perf_counter() - t0)
        nl = h.launches - l0
    got = np.empty(4 * N)
    h.d2h(got, buf)
    ok = np.array_equal(got[:3 * N], np.concatenate([w.ravel() for w in want])) and np.array_equal(got[3 * N:], wu.ravel())
    print(f"seed {seed}: 3 x {n} x {steps} normals + {n} x {steps} uniforms on the device: {best * 1e3:.3f} ms ({nl} launches, "
          f"{used / (3 * N):.5f} generator outputs per normal); NumPy on the host: {t_np * 1e3:.0f} ms; bitwise equal: {ok}", flush=True)
h.free(buf)
p = SVJParams()
for mode in ("device", "host"):
    os.environ["B200MC_REFERENCE_PCG64"] = mode
    e = MonteCarloEngine(p, n, 250, 42, use_sobol=False, use_antithetic=True, use_control_variate=True, rng="reference", handle=h)
    gk = GreeksEngine(p, n, 250, 42, rng="reference", handle=h)
    e.price(22500.0, 22500.0, 1.0)
    t0 = time.perf_counter()
    r = e.price(22500.0, 22500.0, 1.0)
    t1 = time.perf_counter()
    d = gk.delta(22500.0, 22500.0, 1.0)
    v = gk.vega(22500.0, 22500.0, 1.0)
    ga = gk.gamma(22500.0, 22500.0, 1.0)
    t2 = time.perf_counter()
    print(f"rng=reference use_sobol=False front end on the {mode}: price() {1e3 * (t1 - t0):.2f} ms (cached draws), "
          f"delta+vega+gamma {1e3 * (t2 - t1):.2f} ms; price {r['price']:.9f} delta {d['pathwise']:.9f} gamma {ga['gamma']:.6e}", flush=True)
    e2 = MonteCarloEngine(p, n, 250, 43, use_sobol=False, use_antithetic=True, use_control_variate=True, rng="reference", handle=h)
    t0 = time.perf_counter()
    e2.price(22500.0, 22500.0, 1.0)
    print(f"   fresh seed (draws generated inside the call): price() {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
h.close()
