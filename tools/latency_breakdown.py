"""Where a small price() call spends its time: raw C-ABI call vs the Python mirror."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams, _lib  # noqa: E402

p = SVJParams()
h = _lib.default_handle()
sp = _lib.to_params(p)
ks = np.array([22500.0])
out = np.empty((1, _lib.NSUMS))
lib = h.lib
N = 2000


def timeit(f, n=N):
    f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e6


for npaths, steps in ((10_000, 50), (100_000, 50)):
    raw = timeit(lambda: lib.b200mc_price_european(h.h, C.byref(sp), 22500.0, 1.0, steps, npaths, 42, 0, ks.ctypes.data, 1, 1,
                                                    _lib.ANTITHETIC, None, out.ctypes.data))
    wrap = timeit(lambda: h.price_european(p, 22500.0, 1.0, steps, npaths, 42, ks, True, _lib.ANTITHETIC))
    e = MonteCarloEngine(p, npaths, steps, 42, use_sobol=False)
    full = timeit(lambda: e.price(22500.0, 22500.0, 1.0))
    dev = h.malloc(17 * 8)
    asyn = timeit(lambda: lib.b200mc_price_european_async(h.h, C.byref(sp), 22500.0, 1.0, steps, npaths, 42, 0, ks.ctypes.data, 1, 1,
                                                          _lib.ANTITHETIC, None, C.c_void_p(dev)))
    h.synchronize()
    print(f"{npaths} x {steps}: raw C call {raw:.1f} us | Handle.price_european {wrap:.1f} us | MonteCarloEngine.price {full:.1f} us | "
          f"async launch only (throughput) {asyn:.1f} us")
