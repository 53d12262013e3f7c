"""Tail-metric kernel after the move to 16-byte loads: ms per call for 1 / 2 CTAs per SM, float32 / float64, aligned and
misaligned vectors, each checked against the oracle's restatement of engine/risk.py:117-173."""
import os
import sys
import subprocess

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monte_carlo_option_simulator_b200 import _lib  # noqa: E402
from oracle import oracle  # noqa: E402  (checker only)

if len(sys.argv) > 1:                                   # child: one configuration (the env knob is read per call)
    h = _lib.Handle(0)
    for n in (4_000_000, 40_000_000):
        base = np.random.default_rng(0).standard_t(4, size=n + 3) * 0.01
        for dt in (np.float64, np.float32):
            host = base.astype(dt)
            dev = torch.from_numpy(host).cuda()
            for off in (0, 1):
                x = dev[off:off + n]
                got = h.risk_metrics(x.data_ptr(), 0.99, n=n, dtype=dt)
                best = 1e9
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    h.risk_metrics(x.data_ptr(), 0.99, n=n, dtype=dt)
                    e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                line = f"CTAs/SM={os.environ.get('B200MC_RISK_CTAS')} n={n} {np.dtype(dt).name} offset {off}: {best:.3f} ms"
                if off in (0, 1) and n == 4_000_000:
                    w = oracle.risk_metrics(host[off:off + n].astype(np.float64), 0.99)
                    want = [w[k] for k in ("var", "cvar", "skewness", "kurtosis", "excess_kurtosis", "tail_index", "mean", "std")]
                    err = max(abs(g - w) / max(abs(w), 1e-300) for g, w in zip(got, want))
                    line += f"  max rel deviation from the oracle {err:.1e}"
                    assert err < 1e-9, (got, want)
                print(line, flush=True)
    h.close()
else:
    for ctas in ("1", "2"):
        env = dict(os.environ, B200MC_RISK_CTAS=ctas)
        subprocess.run([sys.executable, __file__, "child"], env=env, check=True)
    env = dict(os.environ, B200MC_RISK_TRACE="1")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "risk_probe.py")], env=env, check=False)
