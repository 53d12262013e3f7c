#!/bin/bash
# round 2, call 30: the build that ships (results through the mapped landing buffer by default, leaner Python marshalling):
# full GPU suite, bench line, small-call latencies, calibration wall time (copy vs mapped), smoke().
set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_final.log
timeout 120 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?"
O=gpurun_out/r02_small_call_latency.txt
{ echo "== shipped defaults (B200MC_RESULT unset = mapped)"; timeout 40 python tools/latency_probe.py
  timeout 40 python tools/latency_anatomy.py
  timeout 60 python tools/calibration_timing.py
  echo "== B200MC_RESULT=copy"; B200MC_RESULT=copy timeout 60 python tools/calibration_timing.py
  B200MC_RESULT=copy timeout 40 python tools/latency_probe.py; } > $O 2>&1
cat $O
timeout 60 python __graft_entry__.py smoke > gpurun_out/r02_smoke_final.txt 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke_final.txt
