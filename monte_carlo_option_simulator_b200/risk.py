"""Host mirror of the path-based tail metrics of engine/risk.py (compute_risk_metrics :117-155, Hill :158-173).

The reference sorts the whole P&L vector on the host; here the vector goes to the GPU once (or already lives there)
and the order statistics come from an exact radix select (csrc/risk.cu).  Same keys, same index conventions."""
from __future__ import annotations

import math
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


def _key_to_value(key: int) -> float:
    """Inverse of the order-preserving image used by csrc/risk.cu (negative doubles: ~bits, others: bits | 2^63)."""
    bits = (key & 0x7FFFFFFFFFFFFFFF) if key >> 63 else (~key & 0xFFFFFFFFFFFFFFFF)
    return float(np.array([bits], dtype=np.uint64).view(np.float64)[0])


def compute_risk_metrics_sharded(local_returns, confidence: float = 0.99, *, comm, handle=None) -> Dict[str, float]:
    """compute_risk_metrics (engine/risk.py:117-155) over a vector that is SHARDED across the ranks of `comm`
    (each rank passes its own shard: a NumPy array, or (device_ptr, n, dtype)).  The shards never move: every rank runs
    the radix select on its own keys and only 2 doubles, 8 x 512 counters and 6 doubles are all-reduced, so all ranks
    take the same decisions and return the same global metrics (same index conventions as the reference)."""
    h = handle or _lib.default_handle()
    if isinstance(local_returns, tuple):
        ptr, n_local, dt = local_returns
        s = h.risk_begin(int(ptr), n_local, dt)
    else:
        a = np.asarray(local_returns)
        n_local = a.size
        s = h.risk_begin(a)
    tot = comm.allreduce_sum(np.array([s[0], s[1], float(n_local)]))
    n, m = int(round(tot[2])), int(round(tot[1]))
    if n == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    mean = tot[0] / n                                                        # :137
    cutoff = max(int(n * (1 - confidence)), 0)                               # :128
    rank_c = cutoff if cutoff < n else 0                                     # :129
    want_hill = m > 20                                                       # :150
    k = 0
    if want_hill:
        k = min(max(int(math.sqrt(m)), 10), m - 1)                           # :165-166
    nsel = 2 if want_hill else 1
    prefix, ranks = [0, 0], [rank_c, k]
    for radix_pass in range(7, -1, -1):
        hist = comm.allreduce_sum(h.risk_hist(radix_pass, nsel, prefix).astype(np.float64)).reshape(2, 256)
        for sel in range(nsel):
            cum = 0
            chosen = 255
            for b in range(256):
                c = int(round(hist[sel, b]))
                if cum + c > ranks[sel]:
                    chosen = b
                    break
                cum += c
            ranks[sel] -= cum
            prefix[sel] |= chosen << (8 * radix_pass)
    thr = [_key_to_value(prefix[0]), _key_to_value(prefix[1]) if want_hill else 0.0]
    r2 = comm.allreduce_sum(h.risk_finish(mean, nsel, thr))
    sd = math.sqrt(r2[0] / n)                                                # :138
    sdc = max(sd, 1e-10)                                                     # :141
    skew = (r2[1] / n) / sdc ** 3
    kurt = (r2[2] / n) / sdc ** 4
    if cutoff <= 0:
        cvar = -thr[0]
    elif cutoff >= n:
        cvar = -mean
    else:
        cvar = -(r2[4] + (cutoff - r2[3]) * thr[0]) / cutoff                 # ties at the threshold
    tail = float("nan")
    if want_hill and thr[1] < 0.0 and r2[5] > 0.0:
        tail = k / r2[5]                                                     # :168-173
    return {"var": -thr[0], "cvar": float(cvar), "skewness": float(skew), "kurtosis": float(kurt),
            "excess_kurtosis": float(kurt - 3.0), "tail_index": float(tail), "mean": float(mean), "std": float(sd)}


def terminal_pnl_metrics(params, spot: float, strike: float, T: float, n_paths: int, n_steps: int, seed: int = 42,
                         is_call: bool = True, premium: Optional[float] = None, confidence: float = 0.99, *,
                         handle=None, dtype=np.float64) -> Dict[str, float]:
    """BASELINE config 4 without leaving the GPU: terminal spots (fused simulator) -> discounted option P&L
    D * payoff(S_T) - premium -> compute_risk_metrics.  premium defaults to the Monte Carlo price of the same paths."""
    h = handle or _lib.default_handle()
    dt = np.dtype(dtype)
    n = int(n_paths)
    disc = float(np.exp(-params.r * T))
    S = h.malloc(n * dt.itemsize)
    pnl = h.malloc(n * 8)
    try:
        h.simulate_terminal(params, float(spot), float(T), int(n_steps), n, seed, _lib.FP64 if dt == np.float64 else 0, dt,
                            dev_ptrs=(S, None, None))
        if premium is None:
            h.option_pnl(S, n, strike, is_call, disc, 0.0, pnl, dtype_in=dt)
            premium = float(h.risk_metrics(pnl, confidence, n=n, dtype=np.float64)[6])      # mean of D * payoff
        h.option_pnl(S, n, strike, is_call, disc, premium, pnl, dtype_in=dt)
        out = h.risk_metrics(pnl, confidence, n=n, dtype=np.float64)
    finally:
        h.free(S)
        h.free(pnl)
    res = {k: float(v) for k, v in zip(KEYS, out)}
    res["premium"] = float(premium)
    return res
