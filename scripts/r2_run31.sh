#!/bin/bash
O=gpurun_out/r3f; mkdir -p $O
for n in 4096 512; do
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 24 --warmup 5 --trees-total $n > $O/b_$n.json 2> $O/b_$n.err; echo "$n rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3f/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'us/sim %.2f'%d['us_per_simulation'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'tree in-graph %.2f flushed %.2f'%(r['launch_us_in_graph_no_flush'], r['launch_us']), 'host us/submit %.0f'%d['setup']['host_us_per_submit'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1500:])
P
