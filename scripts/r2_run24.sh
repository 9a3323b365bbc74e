#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 12 $O/pytest.log
timeout 1200 python bench.py --steps 24 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
tail -c 1500 $O/bench_default.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2x/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'us/sim %.2f'%d['us_per_simulation'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'serial %.1fM'%(d['e2e']['serial']['value']/1e6))
        if d.get('selfplay'): print('  selfplay', d['selfplay']['simulations_per_sec'], d['selfplay']['one_engine'])
        e=d['env']; print('  env', e['value'], e['with_torch_policy']['value'], e['saturated']['steps_per_s_by_games_per_gpu'], 'e2e', e['e2e']['value'], e['e2e']['bits_staged']['value'], 'roof', e['roofline']['frac'], e['roofline']['saturated']['frac'])
        if d.get('cpu_baseline'): print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['env_steps_per_s'])
        print('  roof', d['roofline']['frac'], d['roofline']['launch_us'], d['roofline']['launch_us_in_graph_no_flush'], d['roofline']['traffic'], d['roofline']['traffic_source'])
    except Exception as ex:
        print(f, 'ERR', ex)
P
