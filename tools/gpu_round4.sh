#!/bin/bash
mkdir -p gpurun_out
for f in test_ops_gpu test_field_gpu test_render_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "== $f exit $?"; tail -12 gpurun_out/$f.log
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_full.log 2>&1; echo "== bench exit $?"; tail -1 gpurun_out/bench_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --views-per-gpu 8 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "== ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_forward -s 6 -c 2 -o gpurun_out/prof_field2 python bench.py --steps 1 --warmup 1 --views-per-gpu 8 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "== ncu full exit $?"
