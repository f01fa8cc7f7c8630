#!/bin/bash
# 2-GPU checks: NCCL equality test, score / train bench under torchrun
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
nvidia-smi -L
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -n 5
echo "== score N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2k_score2.json 2> gpurun_out/r2k_score2.err; python -c "
import json;d=json.loads(open('gpurun_out/r2k_score2.json').read().strip().splitlines()[-1]);print(round(d['value']/1e6,1),'Mrays/s', round(d['ms_per_step'],1),'ms e2e',round(d['e2e']['value']/1e6,1),d['per_rank_ms_per_step'],d['config']['views_per_gpu'],d['config']['shard'])"; tail -n 4 gpurun_out/r2k_score2.err
echo "== score N=2 contiguous"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --balance contiguous > gpurun_out/r2k_score2c.json 2> gpurun_out/r2k_score2c.err; python -c "
import json;d=json.loads(open('gpurun_out/r2k_score2c.json').read().strip().splitlines()[-1]);print(round(d['value']/1e6,1),'Mrays/s', round(d['ms_per_step'],1),'ms e2e',round(d['e2e']['value']/1e6,1),d['per_rank_ms_per_step'],d['config']['views_per_gpu'],d['config']['shard'])"; tail -n 4 gpurun_out/r2k_score2c.err
echo "== train N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload train --steps 20 --warmup 5 > gpurun_out/r2k_train2.json 2> gpurun_out/r2k_train2.err; tail -c 600 gpurun_out/r2k_train2.json; tail -n 4 gpurun_out/r2k_train2.err
echo "== train N=1"; timeout 600 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/r2k_train1.json 2> gpurun_out/r2k_train1.err; tail -c 600 gpurun_out/r2k_train1.json; tail -n 4 gpurun_out/r2k_train1.err
echo "== train profile"; timeout 600 python tools/train_profile.py > gpurun_out/r2k_train_profile.txt 2>&1; head -n 30 gpurun_out/r2k_train_profile.txt
echo "== round N=1 (100 train steps)"; timeout 900 python bench.py --workload round --steps 2 --warmup 1 --train-steps 100 > gpurun_out/r2k_round1.json 2> gpurun_out/r2k_round1.err; tail -c 900 gpurun_out/r2k_round1.json; tail -n 4 gpurun_out/r2k_round1.err
