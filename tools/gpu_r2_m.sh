#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_render_gpu.py tests/test_golden.py -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -n 3
S='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(round(d["value"]/1e6,2),"M/s",round(d["ms_per_step"],2),"ms e2e",round(d["e2e"]["value"]/1e6,2),d.get("per_rank_ms_per_step"),d["config"].get("views_per_gpu"))'
for K in 2 3 4; do
echo "== score N=1 K=$K"; timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --concurrent-batches $K > gpurun_out/r2m_score1_$K.json 2> gpurun_out/r2m_score1_$K.err; python -c "$S" gpurun_out/r2m_score1_$K.json; tail -n 3 gpurun_out/r2m_score1_$K.err
done
for B in lpt contiguous; do
echo "== score N=2 $B"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 2 --balance $B > gpurun_out/r2m_score2_$B.json 2> gpurun_out/r2m_score2_$B.err; python -c "$S" gpurun_out/r2m_score2_$B.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2m_score2_$B.err | tail -n 3
done
