#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_reference test_ops_gpu test_field_gpu test_render_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "== $f exit $?"; tail -30 gpurun_out/$f.log
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 2 --warmup 1 --views-per-gpu 8 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1; echo "== bench exit $?"; tail -5 gpurun_out/bench_small.log
