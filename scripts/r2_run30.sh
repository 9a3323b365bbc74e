#!/bin/bash
O=gpurun_out/r3e; mkdir -p $O
timeout 300 python scripts/exp_submit_profile.py > $O/submit_profile.txt 2>&1; echo "rc=$?"
grep -v "^$" $O/submit_profile.txt | head -75 | cut -c1-150
