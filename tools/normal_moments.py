"""Statistical certificate of the device generator: moments of N normals against N(0,1), as z-scores.
    python tools/normal_moments.py [n_paths] [n_blocks]      (N = 8 * n_paths * n_blocks draws)"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import _lib  # noqa: E402

n_paths = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000_000
n_blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 32
h = _lib.Handle(0)
for seed in (42, 7, 2026):
    h.timer_begin()
    m = h.normal_moments(seed, n_paths, n_blocks)
    ms = h.timer_end()
    N = m[0]
    mean, m2, m3, m4, cross = (m[k] / N for k in (1, 2, 3, 4, 5))
    # standard errors under N(0,1): mean 1/sqrt N, m2 sqrt(2/N), m3 sqrt(15/N), m4 sqrt(96/N), pair product sqrt(1/(N/2))/... 
    print(f"seed {seed:5d}  N = {N:.3e} draws in {ms:.0f} ms ({N / ms / 1e6:.0f} Gdraws/s)")
    print(f"   mean      {mean:+.3e}            z = {mean * math.sqrt(N):+.2f}")
    print(f"   E z^2 - 1 {m2 - 1:+.3e}            z = {(m2 - 1) / math.sqrt(2 / N):+.2f}")
    print(f"   E z^3     {m3:+.3e}            z = {m3 / math.sqrt(15 / N):+.2f}")
    print(f"   E z^4 - 3 {m4 - 3:+.3e}            z = {(m4 - 3) / math.sqrt(96 / N):+.2f}")
    print(f"   E z_a z_b {2 * cross:+.3e} (pairs)    z = {2 * cross * math.sqrt(N / 2):+.2f}", flush=True)
h.close()
