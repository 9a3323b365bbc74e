#!/bin/bash
mkdir -p gpurun_out/r2e
timeout 1800 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r2e/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e/pytest.log
tail -n 12 gpurun_out/r2e/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e/bench_default.json 2> gpurun_out/r2e/bench_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2e/bench_default.err
HZ_LIB=hanabizero_b200/csrc/libhzb200_trace.so timeout 300 python scripts/exp_env_trace.py > gpurun_out/r2e/env_trace.txt 2>&1
tail -10 gpurun_out/r2e/env_trace.txt
N=512 HZ_LIB=hanabizero_b200/csrc/libhzb200_trace.so timeout 300 python scripts/exp_trace.py > gpurun_out/r2e/trace_512.txt 2>&1
tail -12 gpurun_out/r2e/trace_512.txt
