#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
timeout 1200 python bench.py --steps 12 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
tail -c 800 $O/bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum,launch__shared_mem_per_block_dynamic,launch__shared_mem_per_block_static,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__waves_per_multiprocessor --clock-control none --launch-skip 600 -c 64 --csv --log-file $O/launch_cfg.csv python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --in-flight 1 > $O/ncu_cfg.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'us/sim %.2f'%d['us_per_simulation'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'serial %.1fM'%(d['e2e']['serial']['value']/1e6))
        if d.get('selfplay'): print('  selfplay', d['selfplay']['simulations_per_sec'], d['selfplay']['fraction_of_search_only'])
        if d.get('env'): print('  env', d['env']['value'], d['env']['e2e'])
        if d.get('cpu_baseline'): print('  cpu', d['cpu_baseline']['value'])
    except Exception as e:
        print(f, 'ERR', e)
P
