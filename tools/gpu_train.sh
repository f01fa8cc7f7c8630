#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/test_training_gpu.log 2>&1
echo "== test_training_gpu exit $?"; tail -n 30 gpurun_out/test_training_gpu.log
