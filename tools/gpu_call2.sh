#!/bin/bash
# Round-2 GPU call: golden dump of the new SVJ stream, tests, rates, path-store variants, bench, ncu captures.
# Only small files go to gpurun_out/ (64 MiB cap): the .ncu-rep files stay in /tmp on the box except the fused one.
set -x
mkdir -p gpurun_out
python tests/golden/make_fused_golden.py --dump gpurun_out/fused_draws_r02.npz > gpurun_out/r02_dump.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_fused_modes_against_the_reference_itself > gpurun_out/r02_pytest1.log 2>&1
tail -15 gpurun_out/r02_pytest1.log
timeout 300 python tools/quick_rate.py > gpurun_out/r02_quick_rate.txt 2>&1
timeout 600 python tools/path_store_variants.py > gpurun_out/r02_path_store_variants.txt 2>&1
timeout 300 python tools/numpy_rng_probe.py > gpurun_out/r02_numpy_rng_probe.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err
export NCU_TARGET_REPS=1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:"k_european" -c 4 -o /tmp/r02_fused python tools/ncu_targets.py gbm32 gbm64 heston svj > gpurun_out/r02_ncu_fused.log 2>&1
timeout 900 $NCU -k regex:"k_paths|k_given" -c 3 -o /tmp/r02_store python tools/ncu_targets.py paths32 paths64 given > gpurun_out/r02_ncu_store.log 2>&1
timeout 900 $NCU -k regex:"k_risk|k_hedge|k_qmc|k_cells|k_terminal" -c 40 -o /tmp/r02_callers python tools/ncu_targets.py risk hedge qmc > gpurun_out/r02_ncu_callers.log 2>&1
for r in fused store callers; do
  python tools/ncu_summary.py /tmp/r02_$r.ncu-rep > gpurun_out/r02_ncu_summary_$r.txt 2>&1
done
python tools/ncu_traffic.py gbm_f32_greeks=/tmp/r02_fused.ncu-rep:ILi0ELb0ELb1EfLb1E gbm_f64_greeks=/tmp/r02_fused.ncu-rep:ILi0ELb0ELb1EdLb1E \
   heston_f32_antithetic=/tmp/r02_fused.ncu-rep:ILi2E svj_f32_antithetic=/tmp/r02_fused.ncu-rep:ILi3E \
   paths_f32=/tmp/r02_store.ncu-rep:ffLi paths_f64_out_f32_state=/tmp/r02_store.ncu-rep:fdLi given_normals_svj=/tmp/r02_store.ncu-rep:k_given > gpurun_out/r02_traffic.log 2>&1
cp profiles/r02_ncu_traffic.json gpurun_out/ 2>/dev/null
cp /tmp/r02_fused.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out | tail -30
