#!/bin/bash
# round 2, call 28 (the last GPU minutes of the round): A/B of the fp64 TwoSum build (csrc -DB200MC_FP64_TWOSUM=1,
# libb200mc_twosum.so through B200MC_LIB) against the shipped build, the full GPU suite on the variant (the gate for
# making it the default), then the bench line and smoke() of whichever build ships.
# (build the variant in the container first: make -C monte_carlo_option_simulator_b200/csrc BUILD=build_twosum \
#  OUT=../libb200mc_twosum.so EXTRA=-DB200MC_FP64_TWOSUM=1 -- the .so travels with the snapshot)
set -x
mkdir -p gpurun_out
V=$PWD/monte_carlo_option_simulator_b200/libb200mc_twosum.so
O=gpurun_out/r02_fp64_twosum_ab.txt
{ echo "== shipped build"; timeout 60 python tools/quick_rate.py 10000000 fp64
  echo "== -DB200MC_FP64_TWOSUM=1"; B200MC_LIB=$V timeout 60 python tools/quick_rate.py 10000000 fp64; } > $O 2>&1
cat $O
B200MC_LIB=$V timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_twosum.log 2>&1
echo "pytest (variant) rc=$?" | tee -a $O
tail -5 gpurun_out/r02_pytest_gpu_twosum.log
{ echo "== shipped build, again"; timeout 60 python tools/quick_rate.py 10000000 fp64
  echo "== -DB200MC_FP64_TWOSUM=1, again"; B200MC_LIB=$V timeout 60 python tools/quick_rate.py 10000000 fp64; } >> $O 2>&1
B200MC_LIB=$V timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_twosum.json 2> gpurun_out/r02_bench_n1_twosum.err
echo "bench (variant) rc=$?" | tee -a $O
timeout 100 python __graft_entry__.py smoke > gpurun_out/r02_smoke_final.txt 2>&1
echo "smoke (shipped) rc=$?" | tee -a $O
tail -2 gpurun_out/r02_smoke_final.txt
