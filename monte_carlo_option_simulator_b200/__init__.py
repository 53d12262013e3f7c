"""monte_carlo_option_simulator_b200 -- B200-native Monte Carlo pricing core (libb200mc, sm_100a CUDA) behind the
function signatures of Jay14090/Monte-Carlo-Option-Simulator's engine/monte_carlo.py, engine/greeks.py and the
path-based part of engine/risk.py.  See DESIGN.md and INTEGRATION.md."""
from .models import SVJParams                                   # noqa: F401
from .monte_carlo import (MonteCarloEngine, bs_delta, bs_price,  # noqa: F401
                          _simulate_svj_paths_numba, brownian_bridge_reorder, generate_sobol_normals)
from .greeks import GreeksEngine                                # noqa: F401
from .risk import (HedgingBacktest, LiquidityStress, StressTestEngine,  # noqa: F401
                   compute_risk_metrics)
from .surface import extract_iv_surface, implied_vol           # noqa: F401
from .patch import patch_reference                              # noqa: F401

__all__ = ["SVJParams", "MonteCarloEngine", "GreeksEngine", "compute_risk_metrics", "StressTestEngine", "LiquidityStress",
           "HedgingBacktest", "implied_vol", "extract_iv_surface", "bs_price", "bs_delta", "patch_reference"]
