#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 600 python -m pytest tests/test_training_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -n 8
for cfg in "64 1" "64 2" "64 4" "32 4" "32 8" "128 2"; do set -- $cfg
echo "== bench score vpb $1 concurrent $2"; timeout 900 python bench.py --steps 3 --warmup 2 --views-per-batch $1 --concurrent-batches $2 --no-cpu-baseline > gpurun_out/r2f_$1_$2.json 2> gpurun_out/r2f_$1_$2.err; python -c "
import json;d=json.load(open('gpurun_out/r2f_$1_$2.json'));print(round(d['value']/1e6,1),'Mrays/s', round(d['ms_per_step'],1),'ms e2e',round(d['e2e']['value']/1e6,1),'field Gs/s',round(d['roofline']['gsamples_per_s'],2),'share',round(d['roofline']['kernel_share_of_step'],2),'samples/s',round(d['samples_per_s']/1e9,2))"; tail -n 3 gpurun_out/r2f_$1_$2.err; done
