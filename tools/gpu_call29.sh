#!/bin/bash
# round 2, call 29: A/B of the result hand-over of the synchronous entry points (B200MC_RESULT=copy: device buffer +
# cudaMemcpyAsync, =mapped: the kernel writes the pinned landing buffer itself), then the full GPU suite with "mapped"
# (the gate for making it the default).
set -x
mkdir -p gpurun_out
O=gpurun_out/r02_result_mapped_ab.txt
{ for m in copy mapped copy mapped; do
    echo "== B200MC_RESULT=$m"; B200MC_RESULT=$m timeout 60 python tools/latency_anatomy.py
  done
  for m in copy mapped; do
    echo "== B200MC_RESULT=$m (public API)"; B200MC_RESULT=$m timeout 60 python tools/latency_probe.py
  done; } > $O 2>&1
cat $O
B200MC_RESULT=mapped timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_result_mapped.log 2>&1
echo "pytest (B200MC_RESULT=mapped) rc=$?" | tee -a $O
tail -3 gpurun_out/r02_pytest_gpu_result_mapped.log
