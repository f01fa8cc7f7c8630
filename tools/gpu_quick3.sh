#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/hang_probe.py fused 2>&1 | tail -n 3
for f in test_render_gpu test_edge_cases_gpu; do
  timeout 240 python -m pytest tests/$f.py -m gpu -q --timeout 100 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "== $f exit $?"; tail -n 3 gpurun_out/$f.log
done
timeout 200 python tools/fused_debug.py 2>&1 | tail -n 8 | cut -c1-900
