#!/bin/bash
# round 2, call 20: ncu of the tail-metric kernel on a vector beyond L2 (4e7 float64 values), with the per-instruction page
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:"k_risk_fused" --launch-skip 1 -c 1 -o /tmp/risk40m python tools/ncu_targets.py risk40m > gpurun_out/r02_ncu_risk40m.log 2>&1
tail -3 gpurun_out/r02_ncu_risk40m.log
python tools/ncu_summary.py /tmp/risk40m.ncu-rep > gpurun_out/r02_ncu_k_risk_fused_40M.txt 2>&1
cat gpurun_out/r02_ncu_k_risk_fused_40M.txt | head -60
ncu -i /tmp/risk40m.ncu-rep --page source --csv > gpurun_out/r02_ncu_k_risk_fused_40M_source.csv 2> gpurun_out/r02_ncu_source.err
ls -la gpurun_out/r02_ncu_k_risk_fused_40M_source.csv
du -sh gpurun_out
