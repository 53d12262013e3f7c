#!/bin/bash
# Round-2 GPU call 1: new SVJ stream (golden dump), tests, rates, path-store variants, ncu captures.  Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
python tests/golden/make_fused_golden.py --dump gpurun_out/fused_draws_r02.npz > gpurun_out/r02_dump.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_fused_modes_against_the_reference_itself > gpurun_out/r02_pytest1.log 2>&1
tail -5 gpurun_out/r02_pytest1.log
timeout 300 python tools/quick_rate.py > gpurun_out/r02_quick_rate.txt 2>&1
timeout 600 python tools/path_store_variants.py > gpurun_out/r02_path_store_variants.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err
export NCU_TARGET_REPS=1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:"k_european" -c 4 -o gpurun_out/r02_fused python tools/ncu_targets.py gbm32 gbm64 heston svj > gpurun_out/r02_ncu_fused.log 2>&1
timeout 900 $NCU -k regex:"k_paths|k_given" -c 3 -o gpurun_out/r02_store python tools/ncu_targets.py paths32 paths64 given > gpurun_out/r02_ncu_store.log 2>&1
timeout 900 $NCU -k regex:"k_risk|k_hedge|k_qmc|k_cells|k_terminal" -c 40 -o gpurun_out/r02_callers python tools/ncu_targets.py risk hedge qmc > gpurun_out/r02_ncu_callers.log 2>&1
ls -la gpurun_out | tail -30
