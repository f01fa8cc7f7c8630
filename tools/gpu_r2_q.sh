#!/bin/bash
# Multi-GPU diagnosis (4-GPU box): is the per-rank slowdown at N > 1 host contention, clocks, or the sharding?
#  a) one process alone, 64 views   b) four INDEPENDENT single-GPU processes at once, same command
#  c) torchrun N=4 on the 256-pose batch (contiguous / lpt), with per-rank own times and host-busy times
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
N=${1:-4}
echo "nproc $(nproc)"; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread" ; nvidia-smi -L | wc -l
S='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(round(d["value"]/1e6,2),"M/s",round(d["ms_per_step"],2),"ms own",d["per_rank_ms_per_step"],"host busy",d["host_busy_ms_per_step"]["min"],d["host_busy_ms_per_step"]["max"],"views",d["config"].get("views_per_gpu"),"clk",d["clocks"]["sm_mhz"],d["clocks"]["reasons"])'
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 500 > gpurun_out/r2q_smi.csv 2>&1 &
SMI=$!
echo "== (a) one process alone, 64 views"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 5 --warmup 3 --views 64 --no-cpu-baseline > gpurun_out/r2q_a.json 2> gpurun_out/r2q_a.err; python -c "$S" gpurun_out/r2q_a.json
echo "== (b) $N independent processes at once, 64 views each"
for g in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$g timeout 600 python bench.py --steps 5 --warmup 3 --views 64 --no-cpu-baseline > gpurun_out/r2q_b$g.json 2> gpurun_out/r2q_b$g.err & done
wait $(jobs -p | grep -v $SMI)
for g in $(seq 0 $((N-1))); do python -c "$S" gpurun_out/r2q_b$g.json; done
P=29600
for B in contiguous lpt; do
P=$((P+1)); echo "== (c) torchrun N=$N 256 views $B"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 --balance $B > gpurun_out/r2q_c_$B.json 2> gpurun_out/r2q_c_$B.err; python -c "$S" gpurun_out/r2q_c_$B.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2q_c_$B.err | tail -n 3
done
kill $SMI
python - <<'PY'
import collections
d=collections.defaultdict(list)
for l in open('gpurun_out/r2q_smi.csv'):
    f=[x.strip() for x in l.split(',')]
    if len(f)>=3:
        try: d[f[0]].append((float(f[1].split()[0]), float(f[2].split()[0]), f[3]))
        except Exception: pass
for k,v in sorted(d.items()):
    busy=[x for x in v if x[1]>400]
    if busy: print('gpu',k,'samples under load',len(busy),'sm MHz min/median',min(x[0] for x in busy), sorted(x[0] for x in busy)[len(busy)//2],'power max',max(x[1] for x in busy),'power-cap samples',sum(1 for x in busy if x[2].lower().startswith('active')))
PY
