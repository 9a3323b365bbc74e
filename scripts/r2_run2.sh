#!/bin/bash
mkdir -p gpurun_out/r2b
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest.log
tail -n 25 gpurun_out/r2b/pytest.log
for N in 512 4096; do
  timeout 300 python bench.py --trees $N --steps 10 --warmup 3 --no-cpu-baseline --env-steps 20 > gpurun_out/r2b/bench_$N.json 2> gpurun_out/r2b/bench_$N.err
  N=$N HZ_LIB=hanabizero_b200/csrc/libhzb200_trace.so timeout 300 python scripts/exp_trace.py > gpurun_out/r2b/trace_$N.txt 2>&1
done
timeout 300 python bench.py --trees 2048 --sims 200 --steps 5 --warmup 3 --no-cpu-baseline --env-steps 20 > gpurun_out/r2b/bench_2048x200.json 2> gpurun_out/r2b/bench_2048x200.err
N=512 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 2 -o gpurun_out/r2b/tree512 python scripts/prof_tree.py > gpurun_out/r2b/ncu512.log 2>&1
N=4096 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 2 -o gpurun_out/r2b/tree4096 python scripts/prof_tree.py > gpurun_out/r2b/ncu4096.log 2>&1
ls -la gpurun_out/r2b
