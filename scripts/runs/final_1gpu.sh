#!/bin/bash
O=gpurun_out/r2scale; mkdir -p $O
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > $O/reference_1gpu.json 2> $O/reference_1gpu.err; echo "reference rc=$?"
timeout 1200 python bench.py --steps 20 --warmup 5 > $O/config4_1gpu.json 2> $O/config4_1gpu.err; echo "bench rc=$?"
tail -c 600 $O/config4_1gpu.err
for c in 1 2 3; do
  timeout 600 python bench.py --config $c --steps 24 --warmup 3 --quick --no-cpu-baseline > $O/config${c}_1gpu.json 2> $O/config${c}_1gpu.err; echo "config $c rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2scale/*_1gpu.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.2fM'%(d['value']/1e6), 'one', (d.get('one_search_at_a_time') or {}).get('value'), 'e2e %.2fM'%(d['e2e']['value']/1e6), d['config']['workload'][:60])
        if d.get('selfplay'): print('  selfplay', d['selfplay']['simulations_per_sec'], d['selfplay']['one_engine'])
        if d.get('env') and 'e2e' in d['env']:
            e=d['env']; print('  env', e['value'], e['e2e']['value'])
        if d.get('cpu_baseline'): print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('env_steps_per_s'))
    except Exception as ex:
        print(f, 'ERR', ex)
P
