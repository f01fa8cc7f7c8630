#!/bin/bash
# A/B of an environment knob: field-kernel parity tests + bench line with the knob set, then the default bench line
# usage: tools/gpu_ab.sh APNERF_PAIR_LOADS=1
summ='import json,sys
d=json.loads(sys.stdin.read()); r=d["roofline"]
print("value %.1fM rays/s  e2e %.1fM  ms/step %.2f | field: %.2f Gs/s frac %.3f share %.2f sm_mhz %s"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"],r["gsamples_per_s"],r["frac"],r["kernel_share_of_step"],d["clocks"]["sm_mhz"]))'
echo "== with $1"
env "$1" timeout 300 python -m pytest tests/test_field_gpu.py tests/test_render_gpu.py -m gpu -q --timeout 100 -p no:cacheprovider 2>&1 | tail -n 3
env "$1" timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -n 1 | python -c "$summ"
echo "== default"
timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -n 1 | python -c "$summ"
echo "== with $1 (again)"
env "$1" timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -n 1 | python -c "$summ"
