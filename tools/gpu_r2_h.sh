#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -n 4
echo "== breakdown 64 views"; timeout 600 python tools/kernel_breakdown.py 64 > gpurun_out/r2h_bd64.txt 2>&1; head -n 6 gpurun_out/r2h_bd64.txt; tail -n 4 gpurun_out/r2h_bd64.txt
