#!/bin/bash
# Round-2 evidence: launch list of a short bench, ncu --set full of the repo's two kernels (each program has exited 0
# without ncu first: the same commands had been run plain before), raw pages exported to CSV on the box.
O=gpurun_out/r3c; mkdir -p $O
timeout 300 python scripts/prof_tree.py > $O/plain_tree.log 2>&1; echo "plain tree rc=$?"
timeout 300 python scripts/prof_env.py > $O/plain_env.log 2>&1; echo "plain env rc=$?"
timeout 600 python bench.py --steps 8 --warmup 3 --quick --no-cpu-baseline > $O/plain_bench.json 2> $O/plain_bench.err; echo "plain bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_ncu_launches_bench.csv python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --in-flight 1 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# tree step: caches left as the previous launch left them (as inside a search) ...
N=4096 timeout 600 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 4 -o $O/r02_tree4096_warm python scripts/prof_tree.py > $O/ncu4096w.log 2>&1; echo "ncu tree 4096 warm rc=$?"
# ... and with ncu's default flush before every replay pass (every byte from DRAM)
N=4096 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 4 -o $O/r02_tree4096_cold python scripts/prof_tree.py > $O/ncu4096c.log 2>&1; echo "ncu tree 4096 cold rc=$?"
N=512 timeout 600 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 4 -o $O/r02_tree512_warm python scripts/prof_tree.py > $O/ncu512w.log 2>&1; echo "ncu tree 512 warm rc=$?"
N=4096 timeout 600 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:k_env --launch-skip 40 --launch-count 4 -o $O/r02_env4096_warm python scripts/prof_env.py > $O/ncuenv.log 2>&1; echo "ncu env rc=$?"
N=65536 timeout 600 ncu --set full --clock-control none --cache-control none --kernel-name regex:k_env --launch-skip 40 --launch-count 2 -o $O/r02_env65536_warm python scripts/prof_env.py > $O/ncuenv64k.log 2>&1; echo "ncu env 64k rc=$?"
for f in r02_tree4096_warm r02_tree4096_cold r02_tree512_warm r02_env4096_warm r02_env65536_warm; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.csv 2>/dev/null
done
ls -la $O | head -40
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3c/plain_bench.json').read().strip().splitlines()[-1])
print('value %.1fM'%(d['value']/1e6), 'one %.1fM'%(d['one_search_at_a_time']['value']/1e6), d['setup'])
P
