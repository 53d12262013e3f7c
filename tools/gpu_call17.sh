#!/bin/bash
set -x
mkdir -p gpurun_out
NUMBA_NUM_THREADS=16 timeout 1500 python tools/verify_dropin.py baseline/_ref > gpurun_out/r02_dropin_verify_py_patched.log 2>&1
tail -40 gpurun_out/r02_dropin_verify_py_patched.log
