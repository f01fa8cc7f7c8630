#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 500 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --views-per-gpu 32 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "== ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_forward -s 4 -c 2 -o gpurun_out/prof_field_final python bench.py --steps 1 --warmup 1 --views-per-gpu 16 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "== ncu full exit $?"
