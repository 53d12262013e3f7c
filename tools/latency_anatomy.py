"""Anatomy of a small synchronous price call: kernel duration (CUDA events), launch-only host cost, sync + D2H."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

h = _lib.default_handle()
lib = h.lib
ks = np.array([22500.0])
out = np.empty((1, _lib.NSUMS))
dev = h.malloc(17 * 8)
N = 3000
for name, p in (("svj", SVJParams()), ("gbm", SVJParams.gbm(0.2))):
    sp = _lib.to_params(p)
    for npaths, steps in ((10_000, 50), (100_000, 50), (50_000, 252)):
        args = (h.h, C.byref(sp), 22500.0, 1.0, steps, npaths, 42, 0, ks.ctypes.data, 1, 1, _lib.ANTITHETIC, None)
        lib.b200mc_price_european(*args, out.ctypes.data)
        # (a) kernel duration alone
        h.synchronize(); h.timer_begin()
        for _ in range(50):
            lib.b200mc_price_european_async(*args, C.c_void_p(dev))
        kern_us = h.timer_end() / 50 * 1e3
        # (b) host cost of an async launch when the queue is empty (launch, then wait)
        t = 0.0
        for _ in range(300):
            h.synchronize()
            t0 = time.perf_counter(); lib.b200mc_price_european_async(*args, C.c_void_p(dev)); t += time.perf_counter() - t0
        launch_us = t / 300 * 1e6
        # (c) full synchronous call
        t0 = time.perf_counter()
        for _ in range(N):
            lib.b200mc_price_european(*args, out.ctypes.data)
        full_us = (time.perf_counter() - t0) / N * 1e6
        print(f"{name} {npaths:6d} x {steps:3d}: kernel {kern_us:6.1f} us | host launch call {launch_us:5.1f} us | synchronous call {full_us:6.1f} us "
              f"=> sync + D2H + idle {full_us - launch_us - kern_us:6.1f} us")
