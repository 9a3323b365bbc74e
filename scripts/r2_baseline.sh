#!/bin/bash
# round-2 baseline: where a simulation's time goes as the root batch shrinks (strong-scaling shards)
mkdir -p gpurun_out/r2base
for N in 256 512 1024 2048 4096; do
  HZ_FUSED_CHAIN=1 python scripts/exp_fused_chain.py $N > gpurun_out/r2base/fused_$N.log 2>&1
  N=$N python scripts/exp_chain.py > gpurun_out/r2base/chain_$N.log 2>&1
  python bench.py --trees $N --steps 10 --warmup 3 --no-cpu-baseline --env-steps 20 > gpurun_out/r2base/bench_$N.json 2> gpurun_out/r2base/bench_$N.err
done
HZ_PDL=1 python bench.py --trees 512 --steps 10 --warmup 3 --no-cpu-baseline --env-steps 20 > gpurun_out/r2base/bench_512_pdl.json 2>&1
tail -n 3 gpurun_out/r2base/fused_*.log gpurun_out/r2base/chain_512.log
