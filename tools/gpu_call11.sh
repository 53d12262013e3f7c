#!/bin/bash
set -x
mkdir -p gpurun_out
for v in "" sv3 sv4; do
  if [ -n "$v" ]; then export B200MC_LIB=$PWD/monte_carlo_option_simulator_b200/libb200mc_$v.so; fi
  echo "== variant ${v:-default}" >> gpurun_out/r02_sv_minblocks.txt
  timeout 300 python tools/quick_rate.py 2>&1 | grep -E "heston|svj" >> gpurun_out/r02_sv_minblocks.txt
done
cat gpurun_out/r02_sv_minblocks.txt
