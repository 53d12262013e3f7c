#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest6.log 2>&1
tail -6 gpurun_out/r02_pytest6.log
B200MC_RISK_TRACE=1 timeout 300 python tools/risk_probe.py > gpurun_out/r02_risk_probe_gather2.txt 2>&1
cat gpurun_out/r02_risk_probe_gather2.txt | tail -30
timeout 300 python tools/reference_mode_breakdown.py > gpurun_out/r02_reference_mode_breakdown.txt 2>&1
cat gpurun_out/r02_reference_mode_breakdown.txt
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
