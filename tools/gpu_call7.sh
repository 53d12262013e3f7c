#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest5.log 2>&1
tail -6 gpurun_out/r02_pytest5.log
timeout 300 python tools/risk_probe.py > gpurun_out/r02_risk_probe_gather.txt 2>&1
cat gpurun_out/r02_risk_probe_gather.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err
export NCU_TARGET_REPS=1
timeout 600 ncu --set full --clock-control none -f -k regex:"k_risk" -c 1 -o /tmp/r02_risk2 python tools/ncu_targets.py risk > gpurun_out/r02_ncu_risk2.log 2>&1
python tools/ncu_summary.py /tmp/r02_risk2.ncu-rep > gpurun_out/r02_ncu_summary3_risk.txt 2>&1
du -sh gpurun_out
