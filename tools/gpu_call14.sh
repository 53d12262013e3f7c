#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "path or generate or store or golden or reference_itself" > gpurun_out/r02c_pytest_paths.log 2>&1
tail -4 gpurun_out/r02c_pytest_paths.log
timeout 300 python tools/path_store_probe.py 4000000 > gpurun_out/r02c_path_store_probe.txt 2>&1
cat gpurun_out/r02c_path_store_probe.txt
timeout 600 python tools/path_store_variants.py > gpurun_out/r02c_path_store_variants.txt 2>&1
cat gpurun_out/r02c_path_store_variants.txt
