#!/bin/bash
# Round-2 GPU call 3: full GPU test suite on the regenerated goldens, SVJ with the jump ring, fused tail metrics,
# path-store defaults, bench N=1, ncu captures (summaries only; reps stay on the box).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest2.log 2>&1
tail -15 gpurun_out/r02_pytest2.log
timeout 300 python tools/quick_rate.py > gpurun_out/r02_quick_rate2.txt 2>&1
timeout 300 python tools/risk_probe.py > gpurun_out/r02_risk_probe_fused.txt 2>&1
B200MC_RISK_MULTIKERNEL=1 timeout 300 python tools/risk_probe.py > gpurun_out/r02_risk_probe_multikernel.txt 2>&1
timeout 300 python tools/path_store_probe.py 4000000 > gpurun_out/r02_path_store_probe.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
export NCU_TARGET_REPS=1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:"k_european" -c 1 -o /tmp/r02_svj python tools/ncu_targets.py svj > gpurun_out/r02_ncu_svj.log 2>&1
timeout 900 $NCU -k regex:"k_paths" -c 2 -o /tmp/r02_store python tools/ncu_targets.py paths32 paths64 > gpurun_out/r02_ncu_store.log 2>&1
timeout 900 $NCU -k regex:"k_risk|k_zig|k_scan|k_pcg64" -c 12 -o /tmp/r02_risk python tools/ncu_targets.py risk numpy > gpurun_out/r02_ncu_risk.log 2>&1
for r in svj store risk; do
  python tools/ncu_summary.py /tmp/r02_$r.ncu-rep > gpurun_out/r02_ncu_summary2_$r.txt 2>&1
done
python tools/ncu_traffic.py "svj_f32_antithetic=/tmp/r02_svj.ncu-rep:k_european<3, 1, 0, float, 1>" \
   "paths_f32=/tmp/r02_store.ncu-rep:float, float, 256" "paths_f64_out_f32_state=/tmp/r02_store.ncu-rep:float, double, 256" \
   "risk_fused_f64_4M=/tmp/r02_risk.ncu-rep:k_risk_fused" > gpurun_out/r02_traffic2.log 2>&1
cp profiles/r02_ncu_traffic.json gpurun_out/r02_ncu_traffic.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | tail -30
