#!/bin/bash
# ncu --set full + source of tail-iteration march / composite launches (64 literal-pose views, one batch)
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_march_kernel -s 60 -c 1 -f -o gpurun_out/r2_prof_march_tail python bench.py --steps 1 --warmup 0 --views 64 --views-per-batch 64 --no-cpu-baseline > gpurun_out/r2_prof_march_tail.log 2>&1
tail -n 2 gpurun_out/r2_prof_march_tail.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_composite_kernel -s 60 -c 1 -f -o gpurun_out/r2_prof_comp_tail python bench.py --steps 1 --warmup 0 --views 64 --views-per-batch 64 --no-cpu-baseline > gpurun_out/r2_prof_comp_tail.log 2>&1
tail -n 2 gpurun_out/r2_prof_comp_tail.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_march_kernel -s 0 -c 1 -f -o gpurun_out/r2_prof_march_first python bench.py --steps 1 --warmup 0 --views 64 --views-per-batch 64 --no-cpu-baseline > gpurun_out/r2_prof_march_first.log 2>&1
tail -n 2 gpurun_out/r2_prof_march_first.log
