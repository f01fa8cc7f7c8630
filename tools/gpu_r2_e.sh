#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -n 40
echo "== breakdown 64 views (literal poses)"; timeout 600 python tools/kernel_breakdown.py 64 > gpurun_out/r2e_bd64.txt 2>&1; head -n 12 gpurun_out/r2e_bd64.txt; tail -n 4 gpurun_out/r2e_bd64.txt
