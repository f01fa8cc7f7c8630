#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
for cfg in "148 1" "132 2" "120 2" "112 2" "100 2" "120 3" "112 4"; do set -- $cfg
echo "== bench score field_sms $1 concurrent $2"; APNERF_FIELD_SMS=$1 timeout 900 python bench.py --steps 3 --warmup 2 --views-per-batch 64 --concurrent-batches $2 --no-cpu-baseline > gpurun_out/r2j_$1_$2.json 2> gpurun_out/r2j_$1_$2.err; python -c "
import json;d=json.load(open('gpurun_out/r2j_$1_$2.json'));print(round(d['value']/1e6,1),'Mrays/s', round(d['ms_per_step'],1),'ms e2e',round(d['e2e']['value']/1e6,1),'field Gs/s',round(d['roofline']['gsamples_per_s'],2),'samples/s',round(d['samples_per_s']/1e9,2))"; tail -n 3 gpurun_out/r2j_$1_$2.err; done
