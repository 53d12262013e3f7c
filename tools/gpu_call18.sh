#!/bin/bash
# round 2, call 18: A/B of the cp.async double-buffered given-normals kernel
set -x
mkdir -p gpurun_out
timeout 300 python tools/given_normals_dev_probe.py > gpurun_out/r02_given_async_probe.txt 2>&1
cat gpurun_out/r02_given_async_probe.txt
B200MC_GN_ASYNC=1 timeout 600 python -m pytest tests -m gpu -q -k "given or reference or golden or dropin" 2>&1 | tail -5 > gpurun_out/r02_given_async_pytest.txt
cat gpurun_out/r02_given_async_pytest.txt
B200MC_GN_ASYNC=1 timeout 300 python tools/reference_mode_breakdown.py > gpurun_out/r02_refmode_async.txt 2>&1
tail -15 gpurun_out/r02_refmode_async.txt
