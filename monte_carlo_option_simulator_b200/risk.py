"""Host mirror of the path-based tail metrics of engine/risk.py (compute_risk_metrics :117-155, Hill :158-173).

The reference sorts the whole P&L vector on the host; here the vector goes to the GPU once (or already lives there)
and the order statistics come from an exact radix select (csrc/risk.cu).  Same keys, same index conventions."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import _lib

KEYS = ("var", "cvar", "skewness", "kurtosis", "excess_kurtosis", "tail_index", "mean", "std")


def compute_risk_metrics(returns, confidence: float = 0.99, *, handle=None) -> Dict[str, float]:
    h = handle or _lib.default_handle()
    a = np.asarray(returns)
    if a.size == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")     # what the reference raises (:129)
    out = h.risk_metrics(a, confidence)
    return {k: float(v) for k, v in zip(KEYS, out)}


def terminal_pnl_metrics(params, spot: float, strike: float, T: float, n_paths: int, n_steps: int, seed: int = 42,
                         is_call: bool = True, premium: Optional[float] = None, confidence: float = 0.99, *,
                         handle=None) -> Dict[str, float]:
    """BASELINE config 4: simulate terminal spots on the device, form the discounted option P&L
    D*payoff(S_T) - premium on the host and reduce it with compute_risk_metrics."""
    h = handle or _lib.default_handle()
    S, _, _ = h.simulate_terminal(params, float(spot), float(T), int(n_steps), int(n_paths), seed, 0, np.float64)
    pay = np.maximum(S - strike, 0.0) if is_call else np.maximum(strike - S, 0.0)
    disc = np.exp(-params.r * T)
    if premium is None:
        premium = float(disc * pay.mean())
    return compute_risk_metrics(disc * pay - premium, confidence, handle=h)
