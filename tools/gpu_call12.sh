#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest7.log 2>&1
tail -6 gpurun_out/r02_pytest7.log
timeout 300 python tools/quick_rate.py > gpurun_out/r02_quick_rate4.txt 2>&1
tail -4 gpurun_out/r02_quick_rate4.txt
timeout 300 python tools/cells_probe.py > gpurun_out/r02_cells_probe.txt 2>&1
tail -12 gpurun_out/r02_cells_probe.txt
export NCU_TARGET_REPS=1
timeout 600 ncu --set full --clock-control none -f -k regex:"k_european" -c 1 -o /tmp/r02_svj2 python tools/ncu_targets.py svj > gpurun_out/r02_ncu_svj2.log 2>&1
python tools/ncu_summary.py /tmp/r02_svj2.ncu-rep > gpurun_out/r02_ncu_summary3_svj.txt 2>&1
python tools/ncu_traffic.py "svj_f32_antithetic=/tmp/r02_svj2.ncu-rep:k_european<3, 1, 0, float, 1" > gpurun_out/r02_traffic3.log 2>&1
cp profiles/r02_ncu_traffic.json gpurun_out/r02_ncu_traffic.json
