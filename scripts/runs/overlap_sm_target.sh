#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
for t in 0 74 37 18; do
  N=512 SM_TARGET=$t timeout 300 python scripts/exp_overlap.py > $O/overlap_512_t$t.txt 2>&1; echo "rc=$?"; cat $O/overlap_512_t$t.txt
done
for t in 74 37; do
  N=4096 G=4 SM_TARGET=$t timeout 300 python scripts/exp_overlap.py > $O/overlap_4096_t$t.txt 2>&1; echo "rc=$?"; cat $O/overlap_4096_t$t.txt
done
