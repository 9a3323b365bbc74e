#!/bin/bash
# grouped search: correctness first, then the group-count sweep at the three shard sizes
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests/test_mcts_gpu.py tests/test_search_step_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 5 $O/pytest.log
for g in 1 2 4 8; do
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 10 --warmup 3 --groups $g > $O/bench_4096_g$g.json 2> $O/bench_4096_g$g.err; echo "4096 g$g rc=$?"
done
for g in 1 2 4; do
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 5 --warmup 3 --trees-total 2048 --sims 200 --groups $g > $O/bench_2048x200_g$g.json 2> $O/bench_2048x200_g$g.err; echo "2048x200 g$g rc=$?"
done
for g in 1 2; do
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 10 --warmup 3 --trees-total 512 --groups $g > $O/bench_512_g$g.json 2> $O/bench_512_g$g.err; echo "512 g$g rc=$?"
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 10 --warmup 3 --trees-total 1024 --groups $g > $O/bench_1024_g$g.json 2> $O/bench_1024_g$g.err; echo "1024 g$g rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'us/sim %.2f'%d['us_per_simulation'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'serial %.1fM'%(d['e2e']['serial']['value']/1e6))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-800:])
P
