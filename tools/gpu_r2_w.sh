#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_pipeline_gpu.py tests/test_parity_baseline_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -n 3
S='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(round(d["value"]/1e6,2),"M/s",round(d["ms_per_step"],2),"ms e2e",round(d["e2e"]["value"]/1e6,2))'
for TP in 1 0 1 0; do
echo "== N=1 256 views tail priority $TP"; APNERF_TAIL_PRIORITY=$TP timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_$TP.json 2> gpurun_out/r2w_$TP.err; python -c "$S" gpurun_out/r2w_$TP.json; tail -n 2 gpurun_out/r2w_$TP.err
done
for TP in 1 0; do
echo "== shards x8 on one GPU, tail priority $TP"; PROBE_CHECK_QUICK=1 APNERF_TAIL_PRIORITY=$TP timeout 600 python tools/probe_check.py 8 2>&1 | tail -n 2
done
