#!/bin/bash
# First GPU bring-up: parity suites file by file (each under its own timeout), then diagnostics.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_reference test_ops_gpu; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "== $f exit $?"; tail -25 gpurun_out/$f.log
done
timeout 300 python tools/field_debug.py 2000 > gpurun_out/field_debug.log 2>&1
echo "== field_debug exit $?"; tail -20 gpurun_out/field_debug.log
timeout 600 python -m pytest tests/test_field_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/test_field_gpu.log 2>&1
echo "== test_field_gpu exit $?"; tail -25 gpurun_out/test_field_gpu.log
