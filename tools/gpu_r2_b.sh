#!/bin/bash
# round 2, step B: marcher-written sample points (x01 rows): parity + breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_field_gpu.py tests/test_render_gpu.py tests/test_training_gpu.py tests/test_golden.py tests/test_edge_cases_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -n 5
echo "== breakdown"; timeout 300 python tools/kernel_breakdown.py 32 2>&1 | tail -n 60 > gpurun_out/r2b_new.txt; head -n 12 gpurun_out/r2b_new.txt; tail -n 4 gpurun_out/r2b_new.txt
