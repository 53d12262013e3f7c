#!/bin/bash
set -x
mkdir -p gpurun_out
python tests/golden/make_fused_golden.py --dump gpurun_out/fused_draws_r02b.npz > gpurun_out/r02b_dump.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_fused_modes_against_the_reference_itself > gpurun_out/r02b_pytest.log 2>&1
tail -8 gpurun_out/r02b_pytest.log
timeout 300 python tools/quick_rate.py > gpurun_out/r02b_quick_rate.txt 2>&1
cat gpurun_out/r02b_quick_rate.txt
timeout 600 python tools/rng_quality_probe.py > gpurun_out/r02b_rng_quality_probe.txt 2>&1
cat gpurun_out/r02b_rng_quality_probe.txt
timeout 600 python tools/normal_moments_seeds.py > gpurun_out/r02b_rng_normal_moments_40seeds.txt 2>&1
cat gpurun_out/r02b_rng_normal_moments_40seeds.txt
timeout 600 python tools/bias_check.py 2e9 > gpurun_out/r02b_bias_check_2e9_paths.txt 2>&1
cat gpurun_out/r02b_bias_check_2e9_paths.txt
timeout 300 python tools/path_store_probe.py 4000000 > gpurun_out/r02b_path_store_probe.txt 2>&1
cat gpurun_out/r02b_path_store_probe.txt
