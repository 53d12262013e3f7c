"""One tiny fused launch repeated a few times (for ncu: where does the fixed per-launch time go?)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_carlo_option_simulator_b200 import SVJParams, _lib  # noqa: E402

h = _lib.Handle(0)
p = SVJParams.gbm(0.2)
for i in range(6):
    h.price_european(p, 22500.0, 1.0, 8, 256, 42 + i, [22500.0], True, 0)
print("ok")
