#!/bin/bash
# round 2, call 31 (the last seconds of the GPU budget): price_batch with list strikes + the lean calibration objectives
set -x
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_cells.py tests/test_gpu_parity.py -x -q -m gpu -k "batch or population or grid or smile" > gpurun_out/r02_pytest_gpu_price_batch.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu_price_batch.log
{ timeout 30 python tools/calibration_timing.py; timeout 20 python tools/latency_probe.py; } > gpurun_out/r02_small_call_latency_2.txt 2>&1
cat gpurun_out/r02_small_call_latency_2.txt
