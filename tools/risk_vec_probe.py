"""Tail-metric kernel after the move to 16-byte loads: ms per call for 1 / 2 CTAs per SM, float32 / float64, aligned and
misaligned vectors (parity of exactly these cases: tests/test_gpu_parity.py::test_risk_metrics_device_vectors_at_any_alignment)."""
import os
import sys
import subprocess

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monte_carlo_option_simulator_b200 import _lib  # noqa: E402

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
                line += f"  VaR {got[0]:.9f} CVaR {got[1]:.9f}"
                print(line, flush=True)
    h.close()
else:
    for ctas in ("1", "2"):
        env = dict(os.environ, B200MC_RISK_CTAS=ctas)
        subprocess.run([sys.executable, __file__, "child"], env=env, check=True)
    env = dict(os.environ, B200MC_RISK_TRACE="1")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "risk_probe.py")], env=env, check=False)
