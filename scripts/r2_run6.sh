#!/bin/bash
mkdir -p gpurun_out/r2f
timeout 1800 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r2f/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest.log
tail -n 8 gpurun_out/r2f/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f/bench_default.json 2> gpurun_out/r2f/bench_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2f/bench_default.err
# ncu: launch list of a short bench, then full sets of the two kernels of this repository
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f/r02_ncu_launches_bench.csv python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline > gpurun_out/r2f/ncu_launches.log 2>&1
N=4096 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 4 -o gpurun_out/r2f/r02_tree4096 python scripts/prof_tree.py > gpurun_out/r2f/ncu4096.log 2>&1
N=512 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 4 -o gpurun_out/r2f/r02_tree512 python scripts/prof_tree.py > gpurun_out/r2f/ncu512.log 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name regex:k_env --launch-skip 40 --launch-count 3 -o gpurun_out/r2f/r02_env python scripts/exp_env_trace.py > gpurun_out/r2f/ncu_env.log 2>&1
ls -la gpurun_out/r2f
