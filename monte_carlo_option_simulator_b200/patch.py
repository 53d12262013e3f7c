"""patch_reference(): rebind every import-by-name site of the reference to the CUDA-backed mirrors.

The reference binds the kernel and the engine classes by name in several modules:
engine/greeks.py:16, engine/risk.py:16, engine/calibration.py:19, engine/app.py:26, verify.py:28.  After this call
the FastAPI app, calibration.py, risk.py's StressTestEngine / HedgingBacktest and verify.py run on libb200mc
without any edit to their sources."""
from __future__ import annotations

import importlib
import sys
from typing import List

from . import greeks as _g
from . import monte_carlo as _mc
from . import risk as _r


def patch_reference(package: str = "engine") -> List[str]:
    """Returns the list of 'module.attribute' names that were rebound."""
    done = []

    def rebind(modname, attr, obj):
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                return
        if hasattr(mod, attr):
            setattr(mod, attr, obj)
            done.append(f"{modname}.{attr}")

    mc = f"{package}.monte_carlo"
    for attr in ("_simulate_svj_paths_numba", "MonteCarloEngine", "bs_price", "bs_delta"):
        rebind(mc, attr, getattr(_mc, attr))
    rebind(f"{package}.greeks", "_simulate_svj_paths_numba", _mc._simulate_svj_paths_numba)
    rebind(f"{package}.greeks", "MonteCarloEngine", _mc.MonteCarloEngine)
    rebind(f"{package}.greeks", "GreeksEngine", _g.GreeksEngine)
    rebind(f"{package}.risk", "MonteCarloEngine", _mc.MonteCarloEngine)
    rebind(f"{package}.risk", "compute_risk_metrics", _r.compute_risk_metrics)
    rebind(f"{package}.calibration", "MonteCarloEngine", _mc.MonteCarloEngine)
    for attr, obj in (("MonteCarloEngine", _mc.MonteCarloEngine), ("GreeksEngine", _g.GreeksEngine)):
        rebind(f"{package}.app", attr, obj)
    return done
