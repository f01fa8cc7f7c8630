#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; env $1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.1fM rays/s ms/step %.2f | field share %.2f'%(d['value']/1e6,d['ms_per_step'],r['kernel_share_of_step']))"; }
run "APNERF_SKIP_MIN=1e9 APNERF_MARCH_CFG=256,4"
run "APNERF_SKIP_MIN=32 APNERF_MARCH_CFG=256,4"
run "APNERF_SKIP_MIN=16 APNERF_MARCH_CFG=256,4"
run "APNERF_SKIP_MIN=16 APNERF_MARCH_CFG=128,16"
timeout 600 python -m pytest tests/test_gpu_reference.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -n 2
