#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_callers.py > gpurun_out/r02_multi_gpu_callers_n2.txt 2>&1
tail -12 gpurun_out/r02_multi_gpu_callers_n2.txt
timeout 600 python tools/fuzz_cells.py > gpurun_out/r02_fuzz_cells.txt 2>&1
tail -3 gpurun_out/r02_fuzz_cells.txt
timeout 600 python tools/reference_sobol_probe.py > gpurun_out/r02_reference_sobol_probe.txt 2>&1
tail -8 gpurun_out/r02_reference_sobol_probe.txt
timeout 300 python tools/sanitize_smoke.py > gpurun_out/r02_sanitize_smoke.txt 2>&1
tail -3 gpurun_out/r02_sanitize_smoke.txt
timeout 300 python tools/readme_snippet.py > gpurun_out/r02_readme_snippet.txt 2>&1
tail -12 gpurun_out/r02_readme_snippet.txt
