#!/bin/bash
# ncu --set full of two large field_forward launches (16 views, marching iteration 2 of both members)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_forward -s 4 -c 2 -f -o gpurun_out/r2_prof_field python bench.py --steps 1 --warmup 1 --views-per-gpu 16 --no-cpu-baseline > gpurun_out/r2_prof_field.log 2>&1
tail -n 3 gpurun_out/r2_prof_field.log
