#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 2000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --views-per-gpu 32 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "== ncu list exit $?"; tail -1 gpurun_out/ncu_list.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_composite -s 4 -c 1 -o gpurun_out/prof_composite python bench.py --steps 1 --warmup 1 --views-per-gpu 16 --no-cpu-baseline > gpurun_out/ncu_comp.log 2>&1; echo "== ncu composite exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_march -s 4 -c 1 -o gpurun_out/prof_march python bench.py --steps 1 --warmup 1 --views-per-gpu 16 --no-cpu-baseline > gpurun_out/ncu_march.log 2>&1; echo "== ncu march exit $?"
