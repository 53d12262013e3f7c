#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/rng_quality_probe.py > gpurun_out/r02_rng_quality_probe.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest3.log 2>&1
tail -8 gpurun_out/r02_pytest3.log
timeout 300 python tools/quick_rate.py > gpurun_out/r02_quick_rate3.txt 2>&1
B200MC_RISK_CTAS=1 timeout 300 python tools/risk_probe.py > gpurun_out/r02_risk_probe_ctas1.txt 2>&1
B200MC_RISK_CTAS=2 timeout 300 python tools/risk_probe.py > gpurun_out/r02_risk_probe_ctas2.txt 2>&1
du -sh gpurun_out
