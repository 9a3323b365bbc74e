#!/bin/bash
# 2-GPU sanity of the driver's scaling command with the row-block executor and the deeper pipelines
O=gpurun_out/r2scale_rows; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/config4_2gpu.json 2> $O/config4_2gpu.err; echo "rc=$?"
tail -c 800 $O/config4_2gpu.err | grep -v "^$" | tail -6
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2scale_rows/config4_2gpu.json").read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value %.1fM' % (d['value'] / 1e6), 'one %.1fM' % (d['one_search_at_a_time']['value'] / 1e6), 'e2e %.1fM' % (d['e2e']['value'] / 1e6),
      'weak', (d.get('weak') or {}).get('value'), 'in flight', d['setup']['searches_in_flight'], d['setup']['network_executor'], 'selfplay', d['selfplay']['simulations_per_sec'])
P
