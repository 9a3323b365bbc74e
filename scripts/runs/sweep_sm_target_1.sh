#!/bin/bash
O=gpurun_out/r2v; mkdir -p $O
timeout 600 python -m pytest tests/test_mcts_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
for cfg in "512 8 0" "512 8 16" "512 8 24" "512 4 16" "512 16 16" "1024 8 0" "1024 8 32" "1024 8 48" "1024 4 32" "2048 8 0" "2048 8 64" "2048 4 64" "2048 4 0" "4096 8 0" "4096 4 0"; do
  set -- $cfg
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 32 --warmup 3 --trees-total $1 --in-flight $2 --sm-target $3 > $O/b_$1_d$2_t$3.json 2> $O/b_$1_d$2_t$3.err; echo "$cfg rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2v/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'e2e %.1fM'%(d['e2e']['value']/1e6))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-800:])
P
