#!/bin/bash
# N-GPU A/B of the view scheduler (strong scaling, 256 poses)
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
N=${1:-4}
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -n 3
timeout 300 python tools/probe_check.py 0 2>&1 | head -n 4
S='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(round(d["value"]/1e6,2),"M/s",round(d["ms_per_step"],2),"ms own",{k:round(v,1) for k,v in d["per_rank_ms_per_step"].items()},"probe ms",round(d["config"]["schedule_probe_ms"],2),"views rank0",d["config"].get("views_per_gpu"), "e2e", round(d["e2e"]["value"]/1e6,1))'
P=29700
run() { P=$((P+1)); echo "== N=$N $BAL K=$K $*"; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --balance $BAL --concurrent-batches $K > gpurun_out/r2s.json 2> gpurun_out/r2s.err; python -c "$S" gpurun_out/r2s.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2s.err | tail -n 3; }
BAL=dynamic K=3 run A=1
BAL=dynamic K=3 run APNERF_SHARED_PASSES=5
BAL=dynamic K=4 run APNERF_SHARED_PASSES=12
BAL=lpt K=3 run A=1
