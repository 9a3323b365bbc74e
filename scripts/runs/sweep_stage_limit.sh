#!/bin/bash
O=gpurun_out/r3j; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 6 $O/pytest.log
for cfg in "512 4" "512 2" "1024 4" "1024 2" "2048 4" "4096 4" "4096 2" "4096 0"; do
  set -- $cfg
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 24 --warmup 3 --trees-total $1 --stage-limit $2 > $O/b$1_l$2.json 2> $O/b$1_l$2.err; echo "$cfg rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3j/b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'e2e %.1fM'%(d['e2e']['value']/1e6), 'tree in-graph %.2f flushed %.2f'%(r['launch_us_in_graph_no_flush'], r['launch_us']), d['setup']['tree_stage_limit'], d['setup']['gemm_sm_target'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-800:])
P
