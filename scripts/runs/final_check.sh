#!/bin/bash
# last run of the round: tree step with the in-flight staging limit under ncu, then the whole GPU suite and both bench arms
O=gpurun_out/r3p; mkdir -p $O
N=4096 STAGE_LIMIT=4 timeout 300 python scripts/prof_tree.py > $O/plain_tree.log 2>&1; echo "plain tree rc=$?"
N=4096 STAGE_LIMIT=4 timeout 600 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:k_search_step --launch-skip 30 --launch-count 4 -o $O/r02_tree4096_limit4 python scripts/prof_tree.py > $O/ncu4096l4.log 2>&1; echo "ncu rc=$?"
ncu -i $O/r02_tree4096_limit4.ncu-rep --page raw --csv > $O/r02_tree4096_limit4.csv 2>/dev/null
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 $O/pytest.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
tail -c 400 $O/bench_default.err
python - <<'P'
import json
r=json.loads(open('gpurun_out/r3p/bench_reference.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/r3p/bench_default.json').read().strip().splitlines()[-1])
print('value %.1fM e2e %.1fM one %.1fM reference %.2fM ratio e2e %.1f'%(d['value']/1e6, d['e2e']['value']/1e6, d['one_search_at_a_time']['value']/1e6, r['value']/1e6, d['e2e']['value']/r['value']))
print(d['config']==r['config'], d['plan_vs_module']['root_action_agreement'], d['selfplay']['simulations_per_sec'], d['env']['value'], d['env']['e2e']['value'])
P
