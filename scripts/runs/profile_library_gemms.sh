#!/bin/bash
# ncu --set full of the seven library GEMMs of one recurrent_inference at 4096 and 512 rows (what bounds the searches in flight)
O=gpurun_out/r3r; mkdir -p $O
for n in 4096 512; do
  N=$n timeout 300 python scripts/exp_chain.py > $O/plain_chain_$n.txt 2>&1; echo "plain rc=$?"
  N=$n timeout 600 ncu --set full --clock-control none --cache-control none --kernel-name regex:nvjet --launch-skip 14 --launch-count 7 -o $O/r02_nvjet_$n python scripts/exp_chain.py > $O/ncu_nvjet_$n.log 2>&1; echo "ncu rc=$?"
  ncu -i $O/r02_nvjet_$n.ncu-rep --page raw --csv > $O/r02_nvjet_$n.csv 2>/dev/null
done
ls -la $O
