#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_render_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x --timeout 50 -p no:cacheprovider 2>&1 | tail -n 2
timeout 60 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2z.json 2> gpurun_out/r2z.err; cut -c1-200 gpurun_out/r2z.json; tail -n 2 gpurun_out/r2z.err
