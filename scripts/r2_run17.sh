#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 600 python -m pytest tests/test_env_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 $O/pytest.log
timeout 600 python scripts/exp_env_host.py > $O/env_host.txt 2>&1; echo "env host rc=$?"
cat $O/env_host.txt
