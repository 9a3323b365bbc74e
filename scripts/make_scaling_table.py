"""profiles/r02_scaling.md from the bench lines under gpurun_out/r2scale (scripts/runs/scale.sh, final_1gpu.sh)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "gpurun_out", "r2scale")
def load(f): return json.loads(open(os.path.join(D, f)).read().strip().splitlines()[-1])
rows = [(n, load(f"config4_{n}gpu.json")) for n in (1, 2, 4, 8)]
ref = load("reference_1gpu.json")
out = ["# Round 2: BASELINE config 4 as written — 4096 Hanabi-Full trees x 50 simulations in total, sharded over 1/2/4/8 B200 (strong scaling)",
"",
"`bench.py --gpus N --steps 20 --warmup 5` (one process per GPU, NCCL all-gather of the root statistics per search on a side",
"stream), builder-run on the pool's boxes (`gpurun --gpus N`, `scripts/runs/scale.sh`, `scripts/runs/final_1gpu.sh`); the driver's own",
"SCALE record is the authority.  `value` keeps 7 independent searches in flight per GPU (SearchPipeline: SM-target-sized GEMMs",
"and a 4-node staging limit below 3072 trees per GPU, `r02_sm_target.md`); `one at a time` is the latency figure.",
f"Reference arm on the 1-GPU box's {ref['cpu_baseline']['cores']} host cores (`bench.py --impl reference`): {ref['value']/1e6:.2f} M simulations/s.",
"",
"| GPUs | trees / GPU | value (M sims/s) | x 1 GPU | one at a time (M sims/s) | µs / simulation (one at a time) | e2e host-fed (M sims/s) | weak line: 4096 trees / GPU (M sims/s) | env steps/s at config size (M) | self-play (M sims/s) |",
"|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
v1 = rows[0][1]["value"]
for n, d in rows:
    w = d.get("weak") or {}
    wk = f"{w['value']/1e6:.1f}" if w.get("value") else "(= value)"
    out.append(f"| {n} | {d['config']['trees_per_gpu']} | {d['value']/1e6:.1f} | {d['value']/v1:.2f} | {d['one_search_at_a_time']['value']/1e6:.1f} | {d['us_per_simulation']:.1f} | {d['e2e']['value']/1e6:.1f} | {wk} | {d['env']['value']/1e6:.1f} | {d['selfplay']['simulations_per_sec']/1e6:.1f} |")
out += ["",
"Weak scaling (4096 trees per GPU) stays at 0.95-0.98 of the 1-GPU rate per GPU.",
"",
"Why strong scaling is far from linear — the floor, per launch.  A simulation is a dependent chain of 7 library GEMMs + 1 tree",
"step; none of the eight gets much shorter with fewer rows (`scripts/exp_chain.py`, in-graph, µs):",
"",
"| rows | L1 544->512 | L2 | L3 (+skip) | H1 512->768 | B2 3x(256->256) | A2 (+skip) | B3 3x(256->208) | chain | tree step | simulation (one at a time) |",
"|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|",
"| 4096 | 4.79 | 4.58 | 5.04 | 5.64 | 4.62 | 3.71 | 5.08 | 35.6 | 15.2 | 46.1 |",
"| 2048 | 4.12 | 3.83 | 4.09 | 4.21 | 4.20 | 3.37 | 3.80 | 28.6 | 12.0 | 35.6 |",
"| 1024 | 3.44 | 3.32 | 3.66 | 3.66 | 3.22 | 3.20 | 3.33 | 24.8 | 10.3 | 30.8 |",
"| 512 | 3.23 | 3.14 | 3.41 | 3.52 | 3.25 | 3.12 | 3.06 | 23.3 | 9.9 | 29.1 |",
"| 256 | 3.16 | 3.04 | 3.37 | 3.22 | 2.97 | 3.06 | 2.92 | 22.5 | — | — |",
"",
"An empty kernel in a graph of dependent kernels costs 1.03 µs (`scripts/micro/launch_chain.cu`), so ≈ 8 µs of every",
"simulation is launch dependency alone, and each library GEMM spends ≈ 2.2 µs on its one tile per CTA whatever the row",
"count.  Sharding 4096 trees over 8 GPUs therefore buys 46 -> 29-32 µs per simulation for one search at a time (1.4x);",
"what recovers throughput at small shards is overlap: 7-8 searches in flight per GPU with their GEMMs sized for a share of",
"the SMs and the tree steps small enough to sit next to the GEMM CTAs (17.6 -> 57-61 M simulations/s per GPU at 512 trees on a",
"single GPU; 52 M per GPU in the 8-GPU run, where the slowest rank sets the pace and the per-search all-gather costs ≈ 3 %).",
"Two own tcgen05 executors for the chain were built to attack the per-layer floor and lost to the library",
"(`r02_chain_experiments.md`).",
"",
"Config 5 (16384 trees x 200 simulations, deep-tree stress; `bench.py --config 5 --quick --steps 8`):",
"",
"| GPUs | trees / GPU | value (M sims/s) | one at a time (M sims/s) | ms / search (one at a time) |",
"|---:|---:|---:|---:|---:|"]
for n in (2, 4, 8):
    d = load(f"config5_{n}gpu.json")
    out.append(f"| {n} | {d['config']['trees_per_gpu']} | {d['value']/1e6:.1f} | {d['one_search_at_a_time']['value']/1e6:.1f} | {d['one_search_at_a_time']['ms_per_search']:.2f} |")
out += ["", "Configs 1-3 on one GPU (`bench.py --config C --quick`):", "", "| config | value (M sims/s) | one at a time (M sims/s) | e2e (M sims/s) |", "|---|---:|---:|---:|"]
for c in (1, 2, 3):
    d = load(f"config{c}_1gpu.json")
    name = d["config"]["workload"]
    out.append(f"| {name.split(':')[0]}: {name.split(':')[1].split(',')[0].strip()}, {d['config']['trees_total']} trees | {d['value']/1e6:.2f} | {d['one_search_at_a_time']['value']/1e6:.2f} | {d['e2e']['value']/1e6:.2f} |")
open(os.path.join(ROOT, "profiles", "r02_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[8:15]))
