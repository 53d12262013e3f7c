#!/bin/bash
# final ncu captures of the shipped kernels (summaries only)
set -x
mkdir -p gpurun_out
export NCU_TARGET_REPS=1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:"k_european" -c 4 -o /tmp/f_fused python tools/ncu_targets.py gbm32 gbm64 heston svj > gpurun_out/r02f_ncu_fused.log 2>&1
timeout 900 $NCU -k regex:"k_paths|k_given" -c 3 -o /tmp/f_store python tools/ncu_targets.py paths32 paths64 given > gpurun_out/r02f_ncu_store.log 2>&1
timeout 900 $NCU -k regex:"k_risk|k_zig|k_scan|k_hedge|k_qmc|k_cells" -c 24 -o /tmp/f_callers python tools/ncu_targets.py risk numpy hedge qmc > gpurun_out/r02f_ncu_callers.log 2>&1
for r in fused store callers; do python tools/ncu_summary.py /tmp/f_$r.ncu-rep > gpurun_out/r02f_ncu_summary_$r.txt 2>&1; done
python tools/ncu_traffic.py "gbm_f32_greeks=/tmp/f_fused.ncu-rep:k_european<0, 0, 1, float, 1" "gbm_f64_greeks=/tmp/f_fused.ncu-rep:k_european<0, 0, 1, double, 1" \
  "heston_f32_antithetic=/tmp/f_fused.ncu-rep:k_european<2, 1, 0, float, 1" "svj_f32_antithetic=/tmp/f_fused.ncu-rep:k_european<3, 1, 0, float, 1" \
  "paths_f32=/tmp/f_store.ncu-rep:float, float, 256" "paths_f64_out_f32_state=/tmp/f_store.ncu-rep:float, double, 256" "given_normals_svj=/tmp/f_store.ncu-rep:k_given" \
  "risk_fused_f64_4M=/tmp/f_callers.ncu-rep:k_risk_fused" > gpurun_out/r02f_traffic.log 2>&1
cp profiles/r02_ncu_traffic.json gpurun_out/r02_ncu_traffic.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r02f_ncu_bench.log 2>&1
du -sh gpurun_out
