#!/bin/bash
# round 2, call 19: tail-metric passes with 16-byte loads
set -x
mkdir -p gpurun_out
timeout 600 python tools/risk_vec_probe.py > gpurun_out/r02_risk_vec_probe.txt 2>&1
cat gpurun_out/r02_risk_vec_probe.txt
timeout 600 python -m pytest tests -m gpu -q -k "risk or peer or tail or cfg4 or stress" 2>&1 | tail -5 > gpurun_out/r02_risk_vec_pytest.txt
cat gpurun_out/r02_risk_vec_pytest.txt
