#!/bin/bash
mkdir -p gpurun_out/r2c
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/r2c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
tail -n 30 gpurun_out/r2c/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c/bench_default.json 2> gpurun_out/r2c/bench_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2c/bench_default.err
timeout 300 python bench.py --trees 512 --steps 10 --quick --no-cpu-baseline > gpurun_out/r2c/bench_512.json 2> gpurun_out/r2c/bench_512.err
for N in 512 4096; do
  N=$N HZ_LIB=hanabizero_b200/csrc/libhzb200_trace.so timeout 300 python scripts/exp_trace.py > gpurun_out/r2c/trace_$N.txt 2>&1
done
HZ_LIB=hanabizero_b200/csrc/libhzb200_trace.so timeout 300 python scripts/exp_env_trace.py > gpurun_out/r2c/env_trace.txt 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c/bench_ref.json 2> gpurun_out/r2c/bench_ref.err
ls gpurun_out/r2c
