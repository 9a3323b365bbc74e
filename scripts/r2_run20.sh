#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
N=512 timeout 300 python scripts/exp_overlap.py > $O/overlap_512.txt 2>&1; echo "rc=$?"; cat $O/overlap_512.txt
N=4096 G=4 timeout 300 python scripts/exp_overlap.py > $O/overlap_4096.txt 2>&1; echo "rc=$?"; cat $O/overlap_4096.txt
timeout 600 ncu --metrics gpu__time_duration.sum,launch__shared_mem_per_block_dynamic,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__cluster_dim_x --clock-control none --launch-skip 300 -c 40 --csv --log-file $O/launch_cfg_512.csv python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --in-flight 1 --trees-total 512 > $O/ncu_cfg.log 2>&1; echo "ncu rc=$?"
