#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
timeout 600 python scripts/exp_tune.py > $O/tune.txt 2> $O/tune.err; echo "rc=$?"; cat $O/tune.txt; grep -c candidate $O/tune.err; grep "\->" $O/tune.err | head -20
