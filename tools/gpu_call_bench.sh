#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_call_bench.sh N'   (bench.py only, the driver's flags)
N=${1:-1}
set -x
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final.log 2>&1; tail -3 gpurun_out/r02_pytest_final.log
  python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
  timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
  timeout 300 python tools/risk_probe.py > gpurun_out/r02_risk_probe_final.txt 2>&1
  timeout 300 python tools/quick_rate.py > gpurun_out/r02_quick_rate_final.txt 2>&1
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
fi
tail -c 300 gpurun_out/r02_bench_n$N.err
