#!/bin/bash
O=gpurun_out/r3m; mkdir -p $O
timeout 300 python scripts/prof_tree.py > $O/plain_tree.log 2>&1; echo "plain tree rc=$?"
N=4096 timeout 600 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 4 -o $O/r02_tree4096_warm python scripts/prof_tree.py > $O/ncu4096w.log 2>&1; echo "ncu tree 4096 warm rc=$?"
N=4096 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 4 -o $O/r02_tree4096_cold python scripts/prof_tree.py > $O/ncu4096c.log 2>&1; echo "ncu tree 4096 cold rc=$?"
N=512 timeout 600 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 4 -o $O/r02_tree512_warm python scripts/prof_tree.py > $O/ncu512w.log 2>&1; echo "ncu tree 512 warm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_ncu_launches_bench.csv python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --in-flight 1 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for f in r02_tree4096_warm r02_tree4096_cold r02_tree512_warm; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.csv 2>/dev/null
  ncu -i $O/$f.ncu-rep --page source --csv --print-source sass > $O/$f.sass.csv 2>/dev/null
done
ls -la $O
