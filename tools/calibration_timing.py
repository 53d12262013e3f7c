"""Wall time of the reference's two-stage calibration (engine/calibration.py:195-226: SciPy differential evolution,
200 + 300 iterations, 11 strikes, 20k paths) on the patched engine -- the heaviest caller of small fused launches
(1e4-1e5 objective evaluations of 10-40 us of GPU work each), i.e. the place where the fixed cost per call shows.

    python tools/calibration_timing.py [path of the reference checkout, default baseline/_ref]

Same set-up as the calibration section of tools/verify_dropin.py."""
import logging
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.join(ROOT, "baseline", "_ref")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, ref)
sys.path.insert(0, ROOT)

import engine.calibration  # noqa: E402
from monte_carlo_option_simulator_b200 import patch_reference  # noqa: E402
from monte_carlo_option_simulator_b200.monte_carlo import bs_price  # noqa: E402

logging.getLogger("calibration").setLevel(logging.WARNING)
patch_reference("engine")
ks = np.linspace(0.9, 1.1, 11) * 22500.0
mkt = np.array([bs_price(22500.0, K, 0.08, 0.065, 0.012, 0.16 + 0.4 * (1 - K / 22500.0) ** 2 + 0.2 * max(1 - K / 22500.0, 0), True)
                for K in ks])
for rep in range(2):
    t0 = time.time()
    cal = engine.calibration.CalibrationEngine().calibrate(22500.0, ks, 0.08, mkt, True, num_paths=20_000)
    print(f"[calibration] B200MC_RESULT={os.environ.get('B200MC_RESULT', 'default')}: calibrate (two DE stages, 11 strikes, "
          f"20k paths): {time.time() - t0:.2f} s, stage1 nit={cal['stage1_result']['nit']} nfev={cal['stage1_result'].get('nfev')} "
          f"err={cal['stage1_result']['error']:.4g}, stage2 nit={cal['stage2_result']['nit']} err={cal['stage2_result']['error']:.4g}, "
          f"v0={cal['params'].v0:.4f}", flush=True)
