#!/bin/bash
O=gpurun_out/r3a; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 12 $O/pytest.log
for n in 512 4096; do
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 20 --warmup 5 --trees-total $n > $O/b_$n.json 2> $O/b_$n.err; echo "$n rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3a/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'us/sim %.2f'%d['us_per_simulation'], 'e2e %.1fM'%(d['e2e']['value']/1e6), d['setup']['searches_in_flight'], d['setup']['gemm_sm_target'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1500:])
P
