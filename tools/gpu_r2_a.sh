#!/bin/bash
# round 2, step A: parity of the streamlined field kernel + live breakdown, old (round-1 build) vs new
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_field_gpu.py tests/test_render_gpu.py tests/test_training_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -n 5
echo "== breakdown NEW"; timeout 300 python tools/kernel_breakdown.py 32 2>&1 | tail -n 60 > gpurun_out/r2a_new.txt; head -n 12 gpurun_out/r2a_new.txt; tail -n 1 gpurun_out/r2a_new.txt
echo "== breakdown R1";  APNERF_LIB_PATH=$PWD/build_ab/libapnerf_r1.so timeout 300 python tools/kernel_breakdown.py 32 2>&1 | tail -n 60 > gpurun_out/r2a_r1.txt; head -n 12 gpurun_out/r2a_r1.txt; tail -n 1 gpurun_out/r2a_r1.txt
