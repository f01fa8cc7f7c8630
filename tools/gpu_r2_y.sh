#!/bin/bash
# 8-GPU: draw throttle on (default) / off
mkdir -p gpurun_out
N=${1:-8}
S='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);c=d["config"];print(round(d["value"]/1e6,2),"M/s",round(d["ms_per_step"],2),"ms own",{k:round(v,1) for k,v in d.get("per_rank_ms_per_step",{}).items()},"probe ms",round(c.get("schedule_probe_ms",0),2),"views rank0",c.get("views_per_gpu"),"e2e",round(d["e2e"]["value"]/1e6,2))'
P=29970
run() { P=$((P+1)); echo "== N=$N $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_$1.json 2> gpurun_out/r2y.err; python -c "$S" gpurun_out/r2y_$1.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2y.err | tail -n 3; }
run APNERF_DRAW_THROTTLE=1
run APNERF_DRAW_THROTTLE=0
