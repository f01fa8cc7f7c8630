#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest -m gpu exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 1 gpurun_out/smoke.log | cut -c1-160
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "== ref exit $?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-200
