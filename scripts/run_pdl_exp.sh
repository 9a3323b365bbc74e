#!/bin/bash
# A/B of programmatic dependent launch on the fused tree kernel (HZ_PDL=0/1)
for p in 0 1 0 1; do
  HZ_PDL=$p timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --env-steps 20 2>/dev/null | tail -1 \
    | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('HZ_PDL=$p', round(d['value']/1e6,2), 'Msims/s', round(d['ms_per_step'],4), 'ms')"
done
HZ_PDL=1 timeout 600 python -m pytest tests/test_mcts_gpu.py tests/test_search_step_gpu.py -m gpu -x -q 2>&1 | tail -2
