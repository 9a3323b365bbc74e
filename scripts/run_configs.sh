#!/bin/bash
# BASELINE.json configs on one GPU (config 4 is the default bench line; config 5 = its per-GPU shard at 8 GPUs)
show() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1 |', d['config']['workload'][:60], '|', round(d['value']/1e6,2), 'M sims/s |', round(d['ms_per_step'],3), 'ms/search | e2e', round(d['e2e']['value']/1e6,2), 'M | env', round(d['env']['value']/1e6,1), 'M steps/s | selfplay', round(d['selfplay']['value']/1e3,1), 'k moves/s | tree launch in-graph', round(d['roofline']['launch_us_in_graph_no_flush'],1), 'us')"; }
timeout 300 python bench.py --trees 256 --sims 50 --steps 10 --no-cpu-baseline --env-steps 50 2>/dev/null | show "config2 256x50 global"
timeout 300 python bench.py --trees 1024 --sims 50 --mdp local --stack 4 --steps 10 --no-cpu-baseline --env-steps 50 2>/dev/null | show "config3 1024x50 local stack4"
timeout 300 python bench.py --trees 4096 --sims 50 --steps 10 --no-cpu-baseline --env-steps 50 2>/dev/null | show "config4 4096x50 global"
timeout 600 python bench.py --trees 2048 --sims 200 --steps 5 --no-cpu-baseline --env-steps 50 2>/dev/null | show "config5 shard 2048x200"
timeout 300 python bench.py --trees 4096 --sims 50 --amp none --steps 5 --no-cpu-baseline --env-steps 50 2>/dev/null | show "config4 fp32 network"
