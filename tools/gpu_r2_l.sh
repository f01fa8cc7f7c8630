#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -n 6
S='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(round(d["value"]/1e6,2),"M/s",round(d["ms_per_step"],2),"ms e2e",round(d["e2e"]["value"]/1e6,2),d.get("per_rank_ms_per_step"),d["config"].get("views_per_gpu"))'
echo "== score N=1"; timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2l_score1.json 2> gpurun_out/r2l_score1.err; python -c "$S" gpurun_out/r2l_score1.json; tail -n 3 gpurun_out/r2l_score1.err
echo "== score N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2l_score2.json 2> gpurun_out/r2l_score2.err; python -c "$S" gpurun_out/r2l_score2.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2l_score2.err | tail -n 3
echo "== train N=1"; timeout 600 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/r2l_train1.json 2> gpurun_out/r2l_train1.err; python -c "$S" gpurun_out/r2l_train1.json; tail -n 3 gpurun_out/r2l_train1.err
echo "== train N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --workload train --steps 20 --warmup 5 > gpurun_out/r2l_train2.json 2> gpurun_out/r2l_train2.err; python -c "$S" gpurun_out/r2l_train2.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2l_train2.err | tail -n 3
echo "== round N=1 (100 train steps)"; timeout 900 python bench.py --workload round --steps 2 --warmup 1 --train-steps 100 > gpurun_out/r2l_round1.json 2> gpurun_out/r2l_round1.err; python -c "$S" gpurun_out/r2l_round1.json; tail -n 3 gpurun_out/r2l_round1.err
echo "== round N=1 score only"; timeout 900 python bench.py --workload round --steps 3 --warmup 1 --train-steps 0 > gpurun_out/r2l_round0.json 2> gpurun_out/r2l_round0.err; python -c "$S" gpurun_out/r2l_round0.json; tail -n 3 gpurun_out/r2l_round0.err
