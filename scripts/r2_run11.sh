#!/bin/bash
# 2-GPU strong scaling of config 4 through the pipeline (NCCL gather per search), then the full 1-GPU line
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 --quick --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "2gpu rc=$?"
tail -c 1500 $O/bench_2gpu.err
timeout 1200 python bench.py --steps 12 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
tail -c 800 $O/bench_default.err
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'us/sim %.2f'%d['us_per_simulation'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'serial %.1fM'%(d['e2e']['serial']['value']/1e6))
        if d.get('selfplay'): print('  selfplay', d['selfplay']['simulations_per_sec'], d['selfplay']['fraction_of_search_only'])
        if d.get('env'): print('  env', d['env']['value'], d['env']['e2e']['value'])
        if d.get('cpu_baseline'): print('  cpu', d['cpu_baseline']['value'])
    except Exception as e:
        print(f, 'ERR', e)
P
