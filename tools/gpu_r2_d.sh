#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
echo "== breakdown 64 views (literal poses)"; timeout 600 python tools/kernel_breakdown.py 64 2>&1 | tail -n 100 > gpurun_out/r2d_bd64.txt; head -n 40 gpurun_out/r2d_bd64.txt; tail -n 4 gpurun_out/r2d_bd64.txt
echo "== bench score"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_score.json 2> gpurun_out/r2d_score.err; python -c "
import json;d=json.load(open('gpurun_out/r2d_score.json'));print(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['roofline']['gsamples_per_s'],d['roofline']['frac'],d['roofline']['kernel_share_of_step'],d['config']['mean_samples_per_ray'], d['samples_per_s']/1e9)"; tail -n 5 gpurun_out/r2d_score.err
