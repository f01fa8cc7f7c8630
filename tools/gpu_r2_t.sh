#!/bin/bash
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 900 python -m pytest tests/test_field_gpu.py tests/test_render_gpu.py tests/test_golden.py tests/test_training_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -n 3
for PX in 1 0 1 0; do
echo "== APNERF_FIELD_PAIR_X=$PX"; APNERF_FIELD_PAIR_X=$PX timeout 600 python tools/kernel_breakdown.py 64 > gpurun_out/r2t_bd64_$PX.txt 2>&1; head -n 3 gpurun_out/r2t_bd64_$PX.txt; sed -n 6,16p gpurun_out/r2t_bd64_$PX.txt; tail -n 1 gpurun_out/r2t_bd64_$PX.txt
done
