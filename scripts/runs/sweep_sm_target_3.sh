#!/bin/bash
O=gpurun_out/r3l; mkdir -p $O
for cfg in "2048 8 40" "2048 8 48" "2048 8 56" "2048 8 64" "1024 8 40" "1024 8 48" "1024 8 56" "3072 8 0" "3072 8 64" "3072 8 74" "1536 8 48" "1536 8 56" "1536 8 72"; do
  set -- $cfg
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 24 --warmup 3 --trees-total $1 --in-flight $2 --sm-target $3 > $O/b_$1_d$2_t$3.json 2> $O/b_$1_d$2_t$3.err; echo "$cfg rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3l/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'e2e %.1fM'%(d['e2e']['value']/1e6), d['setup']['searches_in_flight'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-800:])
P
