#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_reference test_ops_gpu test_render_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "== $f exit $?"; tail -n 3 gpurun_out/$f.log
done
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_full.log 2>&1; echo "== bench exit $?"; tail -n 1 gpurun_out/bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.1fM rays/s  e2e %.1fM  ms/step %.2f | field: %.2f Gs/s frac %.3f share %.2f avg_launch %.3f ms'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],r['gsamples_per_s'],r['frac'],r['kernel_share_of_step'],r['avg_launch_ms']))"
