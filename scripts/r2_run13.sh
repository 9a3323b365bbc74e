#!/bin/bash
# cuBLASLt kernel choice under concurrency: heuristic vs tuned alone vs tuned on 4 streams
O=gpurun_out/r2m; mkdir -p $O
for t in 0 1 4; do
  timeout 600 python bench.py --quick --no-cpu-baseline --steps 12 --warmup 3 --tune-streams $t --tune-verbose > $O/bench_4096_t$t.json 2> $O/bench_4096_t$t.err; echo "4096 t$t rc=$?"
done
for t in 0 1 4; do
  timeout 600 python bench.py --quick --no-cpu-baseline --steps 24 --warmup 3 --trees-total 512 --tune-streams $t --tune-verbose > $O/bench_512_t$t.json 2> $O/bench_512_t$t.err; echo "512 t$t rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'us/sim %.2f'%d['us_per_simulation'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'serial %.1fM'%(d['e2e']['serial']['value']/1e6))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1500:])
P
grep "tune: step" $O/bench_4096_t4.err | head -120
