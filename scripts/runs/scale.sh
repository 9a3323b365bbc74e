#!/bin/bash
# usage: scale.sh N   — config 4 (4096 trees in total, strong scaling) full line at N GPUs, config 5 (16384 x 200) quick line
N=$1; O=gpurun_out/r2scale; mkdir -p $O
run() { if [ "$N" = "1" ]; then python bench.py "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py "$@"; fi; }
timeout 900 bash -c "$(declare -f run); N=$N; run --gpus $N --steps 20 --warmup 5 --no-cpu-baseline" > $O/config4_${N}gpu.json 2> $O/config4_${N}gpu.err; echo "config4 x$N rc=$?"
tail -c 600 $O/config4_${N}gpu.err | grep -v "^$" | tail -5
timeout 900 bash -c "$(declare -f run); N=$N; run --gpus $N --steps 8 --warmup 3 --config 5 --quick --no-cpu-baseline" > $O/config5_${N}gpu.json 2> $O/config5_${N}gpu.err; echo "config5 x$N rc=$?"
tail -c 600 $O/config5_${N}gpu.err | grep -v "^$" | tail -5
python - <<'P'
import json,glob,sys
for f in sorted(glob.glob('gpurun_out/r2scale/*gpu.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'n_gpus', d['n_gpus'], 'value %.1fM'%(d['value']/1e6), 'one-at-a-time %.1fM'%(d['one_search_at_a_time']['value']/1e6), 'us/sim %.2f'%d['us_per_simulation'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'weak', (d.get('weak') or {}).get('value'), 'env', (d.get('env') or {}).get('value'))
    except Exception as ex:
        print(f, 'ERR', ex)
P
