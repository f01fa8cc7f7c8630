#!/bin/bash
# round 2, step C: full GPU suite + new bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -n 6
echo "== bench score"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_score.json 2> gpurun_out/r2c_score.err; tail -c 3000 gpurun_out/r2c_score.json; tail -n 5 gpurun_out/r2c_score.err
echo "== bench score vpb 256"; timeout 900 python bench.py --steps 5 --warmup 3 --views-per-batch 256 --no-cpu-baseline > gpurun_out/r2c_score256.json 2> gpurun_out/r2c_score256.err; python -c "
import json;d=json.load(open('gpurun_out/r2c_score256.json'));print(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['roofline']['gsamples_per_s'],d['roofline']['frac'],d['roofline']['kernel_share_of_step'],d['config']['mean_samples_per_ray'])"; tail -n 5 gpurun_out/r2c_score256.err
echo "== bench train"; timeout 600 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/r2c_train.json 2> gpurun_out/r2c_train.err; tail -c 1500 gpurun_out/r2c_train.json; tail -n 5 gpurun_out/r2c_train.err
