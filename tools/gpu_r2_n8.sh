#!/bin/bash
# 8-GPU (or N-GPU: pass N) scaling checks
N=${1:-8}
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
nvidia-smi -L | wc -l
S='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(round(d["value"]/1e6,2),"M/s",round(d["ms_per_step"],2),"ms e2e",round(d["e2e"]["value"]/1e6,2),d.get("per_rank_ms_per_step"),d["config"].get("views_per_gpu"))'
P=29530
for B in contiguous lpt dynamic; do
P=$((P+1)); echo "== score N=$N $B"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 --balance $B > gpurun_out/r2n_score${N}_$B.json 2> gpurun_out/r2n_score${N}_$B.err; python -c "$S" gpurun_out/r2n_score${N}_$B.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2n_score${N}_$B.err | tail -n 3
done
P=$((P+1)); echo "== score N=$N weak 64/gpu contiguous"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 3 --warmup 2 --scaling weak --views-per-gpu 128 --balance contiguous > gpurun_out/r2n_score${N}_weak.json 2> gpurun_out/r2n_score${N}_weak.err; python -c "$S" gpurun_out/r2n_score${N}_weak.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2n_score${N}_weak.err | tail -n 3
P=$((P+1)); echo "== train N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload train --steps 20 --warmup 5 > gpurun_out/r2n_train$N.json 2> gpurun_out/r2n_train$N.err; python -c "$S" gpurun_out/r2n_train$N.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2n_train$N.err | tail -n 3
P=$((P+1)); echo "== round N=$N (200 train steps)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload round --steps 2 --warmup 1 --train-steps 200 > gpurun_out/r2n_round$N.json 2> gpurun_out/r2n_round$N.err; python -c "$S" gpurun_out/r2n_round$N.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2n_round$N.err | tail -n 3
