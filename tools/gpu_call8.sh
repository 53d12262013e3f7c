#!/bin/bash
set -x
mkdir -p gpurun_out
B200MC_RISK_TRACE=1 timeout 300 python - > gpurun_out/r02_risk_trace.txt 2>&1 <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from monte_carlo_option_simulator_b200 import _lib
h = _lib.Handle(0)
for n in (4_000_000, 40_000_000):
    x = torch.from_numpy(np.random.default_rng(0).standard_t(4, size=n) * 0.01).cuda()
    for r in range(3):
        h.risk_metrics(x.data_ptr(), 0.99, n=n, dtype=np.float64)
    x32 = x.float()
    for r in range(2):
        h.risk_metrics(x32.data_ptr(), 0.99, n=n, dtype=np.float32)
h.close()
PY
cat gpurun_out/r02_risk_trace.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err
tail -c 300 gpurun_out/r02_bench_n1_c.err
timeout 200 python tools/pipe_probe.py > gpurun_out/r02_pipe_probe.txt 2>&1
