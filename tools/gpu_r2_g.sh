#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -n 4
echo "== breakdown 64 views"; timeout 600 python tools/kernel_breakdown.py 64 > gpurun_out/r2g_bd64.txt 2>&1; head -n 6 gpurun_out/r2g_bd64.txt; tail -n 4 gpurun_out/r2g_bd64.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_march_kernel -s 60 -c 1 -f -o gpurun_out/r2g_prof_march_tail python bench.py --steps 1 --warmup 0 --views 64 --views-per-batch 64 --no-cpu-baseline > gpurun_out/r2g_prof_march_tail.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_composite_kernel -s 60 -c 1 -f -o gpurun_out/r2g_prof_comp_tail python bench.py --steps 1 --warmup 0 --views 64 --views-per-batch 64 --no-cpu-baseline > gpurun_out/r2g_prof_comp_tail.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
