#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
APNERF_PROBE_MIN_SAMPLES=16 timeout 600 python tools/probe_check.py 8 2>&1 | tee gpurun_out/r2r_probe8.txt
APNERF_PROBE_RAYS=100 APNERF_PROBE_MIN_SAMPLES=16 timeout 600 python tools/probe_check.py 4 2>&1 | tee gpurun_out/r2r_probe4.txt
