#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
S='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(round(d["value"]/1e6,2),"M/s",round(d["ms_per_step"],2),"ms e2e",round(d["e2e"]["value"]/1e6,2),d["config"].get("views_per_gpu"))'
for cfg in "natural consecutive 3" "heavy consecutive 3" "heavy strided 3" "heavy consecutive 4" "natural consecutive 4"; do set -- $cfg
echo "== N=1 256 views order=$1 batching=$2 K=$3"; APNERF_ORDER=$1 APNERF_BATCHING=$2 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --concurrent-batches $3 > gpurun_out/r2o.json 2> gpurun_out/r2o.err; python -c "$S" gpurun_out/r2o.json; tail -n 2 gpurun_out/r2o.err
done
for cfg in "natural consecutive 32 3" "heavy consecutive 16 3" "heavy consecutive 11 3" "heavy consecutive 8 4" "natural consecutive 11 3" "heavy strided 11 3"; do set -- $cfg
echo "== N=1 first 32 views order=$1 batching=$2 vpb=$3 K=$4"; APNERF_ORDER=$1 APNERF_BATCHING=$2 timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --views 32 --views-per-batch $3 --concurrent-batches $4 > gpurun_out/r2o.json 2> gpurun_out/r2o.err; python -c "$S" gpurun_out/r2o.json; tail -n 2 gpurun_out/r2o.err
done
