#!/bin/bash
# Round-end profiling pass (see profiles/): launch list of one bench step + ncu --set full of the hot kernels.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 500 --csv --log-file gpurun_out/launches_final.csv $B --views-per-gpu 32 > gpurun_out/ncu_list.log 2>&1; echo "== ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_forward -s 4 -c 2 -o gpurun_out/prof_field_final $B --views-per-gpu 16 > gpurun_out/ncu_full.log 2>&1; echo "== ncu field exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_composite -s 4 -c 1 -o gpurun_out/prof_composite_final $B --views-per-gpu 16 > gpurun_out/ncu_comp.log 2>&1; echo "== ncu composite exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_march -s 4 -c 1 -o gpurun_out/prof_march_final $B --views-per-gpu 16 > gpurun_out/ncu_march2.log 2>&1; echo "== ncu march exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:field_backward|field_forward|hashgrid_encode_bwd" -s 9 -c 3 -o gpurun_out/prof_train_final python tools/train_grad_check.py > gpurun_out/ncu_train.log 2>&1; echo "== ncu train exit $?"
