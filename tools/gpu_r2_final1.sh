#!/bin/bash
# Final 1-GPU pass of the round: full GPU test suite, smoke, default bench, reference arm
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; echo "== pytest -m gpu exit $?"; tail -n 3 gpurun_out/r2f_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "== smoke exit $?"; tail -n 1 gpurun_out/r2f_smoke.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "== bench exit $?"; cut -c1-330 gpurun_out/r2f_bench.json; tail -n 3 gpurun_out/r2f_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "== ref exit $?"; cut -c1-260 gpurun_out/r2f_ref.json
