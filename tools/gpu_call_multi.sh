#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_call_multi.sh N'
N=${1:-2}
set -x
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest4.log 2>&1; tail -5 gpurun_out/r02_pytest4.log; python tools/risk_probe.py > gpurun_out/r02_risk_probe_final.txt 2>&1; fi
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 600 gpurun_out/r02_bench_n$N.err
timeout 600 $RUN tools/cfg4_multi_gpu.py > gpurun_out/r02_cfg4_multi_gpu_n$N.txt 2>&1
tail -3 gpurun_out/r02_cfg4_multi_gpu_n$N.txt
timeout 300 $RUN tools/peer_allreduce_check.py > gpurun_out/r02_peer_allreduce_n$N.txt 2>&1
tail -3 gpurun_out/r02_peer_allreduce_n$N.txt
if [ "$N" = "2" ]; then
  B200MC_EXCHANGE=nccl timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/r02_bench_n${N}_nccl.json 2> gpurun_out/r02_bench_n${N}_nccl.err
  timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
fi
du -sh gpurun_out
