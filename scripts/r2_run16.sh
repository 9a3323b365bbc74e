#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
timeout 600 python scripts/exp_env_host.py > $O/env_host.txt 2>&1; echo "env host rc=$?"
cat $O/env_host.txt
