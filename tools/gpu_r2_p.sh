#!/bin/bash
# Round-2 evidence pass: GPU tests, smoke, default bench + reference arm, live kernel breakdown, ncu launch list of a
# bench step and ncu --set full of the field kernel (three large launches) -> profiles/r02_*
mkdir -p gpurun_out
python active-perception-using-neural-radiance-fields_b200/csrc/build.py > /dev/null
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r2p_pytest.log 2>&1; echo "== pytest -m gpu exit $?"; tail -n 3 gpurun_out/r2p_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2p_smoke.log 2>&1; echo "== smoke exit $?"; tail -n 1 gpurun_out/r2p_smoke.log | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "== bench exit $?"; cut -c1-400 gpurun_out/r2p_bench.json; tail -n 3 gpurun_out/r2p_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2p_ref.json 2> gpurun_out/r2p_ref.err; echo "== ref exit $?"; cut -c1-300 gpurun_out/r2p_ref.json
timeout 600 python tools/kernel_breakdown.py 64 > gpurun_out/r2p_bd64.txt 2>&1; head -n 6 gpurun_out/r2p_bd64.txt; tail -n 4 gpurun_out/r2p_bd64.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2p_launches_v64.csv python bench.py --steps 1 --warmup 1 --views 64 --no-cpu-baseline > gpurun_out/r2p_ncu_list.log 2>&1; echo "== ncu list exit $?"; python tools/summarize_launches.py gpurun_out/r2p_launches_v64.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_forward -s 1 -c 3 -f -o gpurun_out/r2p_prof_field python tools/field_profile_target.py 16 > gpurun_out/r2p_prof_field.log 2>&1; echo "== ncu field exit $?"; tail -n 2 gpurun_out/r2p_prof_field.log
ls -la gpurun_out/r2p_*
