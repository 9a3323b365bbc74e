#!/bin/bash
O=gpurun_out/r3h; mkdir -p $O
for rep in 1 2; do for cap in 0 4; do
  HZ_EXP_CAP=$cap timeout 300 python bench.py --quick --no-cpu-baseline --steps 8 --warmup 3 --trees-total 2048 --sims 200 > $O/b200_cap${cap}_$rep.json 2> $O/b200_cap${cap}_$rep.err; echo "$cap rc=$?"
  HZ_EXP_CAP=$cap timeout 300 python bench.py --quick --no-cpu-baseline --steps 24 --warmup 3 --trees-total 2048 > $O/b50_cap${cap}_$rep.json 2> $O/b50_cap${cap}_$rep.err; echo "$cap rc=$?"
done; done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3h/b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'e2e %.1fM'%(d['e2e']['value']/1e6), 'tree in-graph %.2f flushed %.2f'%(r['launch_us_in_graph_no_flush'], r['launch_us']))
    except Exception as e:
        print(f, 'ERR', e)
P
