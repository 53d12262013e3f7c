#!/bin/bash
# round 2, call 21: does FFMA2 free issue slots?  + alignment test of the tail metrics
set -x
mkdir -p gpurun_out
timeout 600 python tools/pipe_probe.py > gpurun_out/r02_pipe_probe_ffma2.txt 2>&1
tail -12 gpurun_out/r02_pipe_probe_ffma2.txt
timeout 600 python -m pytest tests -m gpu -q -k "risk_metrics" 2>&1 | tail -4
