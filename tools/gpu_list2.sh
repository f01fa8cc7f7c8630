#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --views-per-gpu 32 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "== ncu list exit $?"
