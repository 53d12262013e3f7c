#!/bin/bash
# round 2, call 27: A/B of the three-sweep (ILP) tile walk of the given-normals kernel
set -x
mkdir -p gpurun_out
timeout 300 python tools/given_normals_dev_probe.py > gpurun_out/r02_given_ilp_probe.txt 2>&1
cat gpurun_out/r02_given_ilp_probe.txt
