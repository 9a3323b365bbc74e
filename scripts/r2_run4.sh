#!/bin/bash
mkdir -p gpurun_out/r2d
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest.log
tail -n 30 gpurun_out/r2d/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d/bench_default.json 2> gpurun_out/r2d/bench_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2d/bench_default.err
HZ_LIB=hanabizero_b200/csrc/libhzb200_trace.so timeout 300 python scripts/exp_env_trace.py > gpurun_out/r2d/env_trace.txt 2>&1
tail -12 gpurun_out/r2d/env_trace.txt
