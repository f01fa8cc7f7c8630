#!/bin/bash
# quick GPU regression: renderer / field / edge tests + the bench line, summarised
timeout 300 python -m pytest tests/test_render_gpu.py tests/test_edge_cases_gpu.py -m gpu -q --timeout 100 -p no:cacheprovider 2>&1 | tail -n 3
timeout 400 python bench.py --no-cpu-baseline 2>&1 | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.1fM rays/s  e2e %.1fM  ms/step %.2f | field: %.2f Gs/s frac %.3f share %.2f launches %d'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],r['gsamples_per_s'],r['frac'],r['kernel_share_of_step'],d['gpu_launches']))"
