#!/bin/bash
# row-block resident network executor: parity + timings, then the search bench with it
O=gpurun_out/rc4; mkdir -p $O
timeout 240 python scripts/exp_rowchain.py > $O/exp.log 2>&1; echo "exp rc=$?"; tail -n 45 $O/exp.log
timeout 300 python bench.py --quick --no-cpu-baseline --executor rows > $O/bench_rows.json 2> $O/bench_rows.err; echo "bench rows rc=$?"
tail -n 5 $O/bench_rows.err
python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/rc4/bench_rows.json"))
    print("rows: value %.1fM e2e %.1fM one %.1fM" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["one_search_at_a_time"]["value"] / 1e6), d.get("plan_vs_module"))
except Exception as e:
    print("no bench line", e)
P
