#!/bin/bash
O=gpurun_out/final; mkdir -p $O
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
tail -c 300 $O/bench_default.err
python - <<'P'
import json
r=json.loads(open('gpurun_out/final/bench_reference.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/final/bench_default.json').read().strip().splitlines()[-1])
print('value %.1fM e2e %.1fM one %.1fM reference %.2fM ratio e2e %.1f'%(d['value']/1e6, d['e2e']['value']/1e6, d['one_search_at_a_time']['value']/1e6, r['value']/1e6, d['e2e']['value']/r['value']))
print(d['config']==r['config'], d['plan_vs_module']['root_action_agreement'], d['selfplay']['simulations_per_sec'], d['env']['value'], d['env']['e2e']['value'], d['env']['roofline']['traffic'], d['roofline']['frac'], d['roofline']['traffic'])
P
