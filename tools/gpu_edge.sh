#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_edge_cases_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/test_edge.log 2>&1
echo "== edge exit $?"; tail -n 30 gpurun_out/test_edge.log
