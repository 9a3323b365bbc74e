#!/bin/bash
# round 2, run 1: parity of the staged tree kernel + first timings
mkdir -p gpurun_out/r2a
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
tail -n 15 gpurun_out/r2a/pytest.log
./scripts/micro/launch_chain > gpurun_out/r2a/launch_chain.txt 2>&1
for N in 512 4096; do
  timeout 300 python bench.py --trees $N --steps 10 --warmup 3 --no-cpu-baseline --env-steps 20 > gpurun_out/r2a/bench_$N.json 2> gpurun_out/r2a/bench_$N.err
  N=$N HZ_LIB=hanabizero_b200/csrc/libhzb200_trace.so timeout 300 python scripts/exp_trace.py > gpurun_out/r2a/trace_$N.txt 2>&1
done
timeout 300 python bench.py --trees 2048 --sims 200 --steps 5 --warmup 3 --no-cpu-baseline --env-steps 20 > gpurun_out/r2a/bench_2048x200.json 2> gpurun_out/r2a/bench_2048x200.err
cat gpurun_out/r2a/launch_chain.txt
