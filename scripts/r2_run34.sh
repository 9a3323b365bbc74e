#!/bin/bash
O=gpurun_out/r3i; mkdir -p $O
for n in 512 1024; do for cap in 0 4 8; do
  HZ_EXP_CAP=$cap timeout 300 python bench.py --quick --no-cpu-baseline --steps 24 --warmup 3 --trees-total $n > $O/b${n}_cap${cap}.json 2> $O/b${n}_cap${cap}.err; echo "$n $cap rc=$?"
done; done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3i/b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'e2e %.1fM'%(d['e2e']['value']/1e6), 'tree in-graph %.2f flushed %.2f'%(r['launch_us_in_graph_no_flush'], r['launch_us']))
    except Exception as e:
        print(f, 'ERR', e)
P
