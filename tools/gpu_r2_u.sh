#!/bin/bash
# 8-GPU validation of the defaults: score (dynamic scheduler) vs contiguous, training step, full round
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
N=${1:-8}
echo "nproc $(nproc), gpus $(nvidia-smi -L | wc -l)"
S='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);c=d["config"];print(round(d["value"]/1e6,2),"M/s" if d["unit"]=="rays/s" else d["unit"],round(d["ms_per_step"],2),"ms own",{k:round(v,1) for k,v in d.get("per_rank_ms_per_step",{}).items()},"probe ms",round(c.get("schedule_probe_ms",0),2),"views rank0",c.get("views_per_gpu"),"e2e",round(d["e2e"]["value"]/1e6,2))'
P=29800
P=$((P+1)); echo "== score N=$N default"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2u_score$N.json 2> gpurun_out/r2u_score$N.err; python -c "$S" gpurun_out/r2u_score$N.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2u_score$N.err | tail -n 3
P=$((P+1)); echo "== score N=$N contiguous"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 10 --warmup 3 --balance contiguous > gpurun_out/r2u_score${N}c.json 2> gpurun_out/r2u_score${N}c.err; python -c "$S" gpurun_out/r2u_score${N}c.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2u_score${N}c.err | tail -n 3
P=$((P+1)); echo "== train N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload train --steps 30 --warmup 5 > gpurun_out/r2u_train$N.json 2> gpurun_out/r2u_train$N.err; python -c "$S" gpurun_out/r2u_train$N.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2u_train$N.err | tail -n 3
P=$((P+1)); echo "== round N=$N (200 train steps)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload round --steps 2 --warmup 1 --train-steps 200 > gpurun_out/r2u_round$N.json 2> gpurun_out/r2u_round$N.err; python -c "$S" gpurun_out/r2u_round$N.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2u_round$N.err | tail -n 3
