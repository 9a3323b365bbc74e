#!/bin/bash
# searches in flight on their own streams: parity suite, then the depth sweep at the config-4 shard sizes
O=gpurun_out/r2j; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 5 $O/pytest.log
for d in 2 3 4 6; do
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 12 --warmup 3 --in-flight $d > $O/bench_4096_d$d.json 2> $O/bench_4096_d$d.err; echo "4096 d$d rc=$?"
done
for d in 2 3 4 8; do
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 24 --warmup 3 --trees-total 512 --in-flight $d > $O/bench_512_d$d.json 2> $O/bench_512_d$d.err; echo "512 d$d rc=$?"
done
for d in 3 4; do
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 6 --warmup 3 --trees-total 2048 --sims 200 --in-flight $d > $O/bench_2048x200_d$d.json 2> $O/bench_2048x200_d$d.err; echo "2048x200 d$d rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'us/sim %.2f'%d['us_per_simulation'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'serial %.1fM'%(d['e2e']['serial']['value']/1e6))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1500:])
P
