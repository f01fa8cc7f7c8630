#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -n 4
for cfg in "64 1 12" "64 2 12" "64 3 12" "64 4 12" "64 3 6" "64 3 24" "32 4 12" "32 6 12"; do set -- $cfg
echo "== bench score vpb $1 concurrent $2 stagger $3"; APNERF_STAGGER=$3 timeout 900 python bench.py --steps 3 --warmup 2 --views-per-batch $1 --concurrent-batches $2 --no-cpu-baseline > gpurun_out/r2i_$1_$2_$3.json 2> gpurun_out/r2i_$1_$2_$3.err; python -c "
import json;d=json.load(open('gpurun_out/r2i_$1_$2_$3.json'));print(round(d['value']/1e6,1),'Mrays/s', round(d['ms_per_step'],1),'ms e2e',round(d['e2e']['value']/1e6,1),'field Gs/s',round(d['roofline']['gsamples_per_s'],2),'samples/s',round(d['samples_per_s']/1e9,2))"; tail -n 3 gpurun_out/r2i_$1_$2_$3.err; done
