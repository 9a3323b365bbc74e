#!/bin/bash
# searches in flight / tree staging with the row-block executor
O=gpurun_out/rcsweep; mkdir -p $O
for cfg in "5 4" "10 4" "15 4" "10 0" "6 4"; do
  set -- $cfg
  timeout 200 python bench.py --quick --no-cpu-baseline --executor rows --steps 30 --in-flight $1 --stage-limit $2 > $O/b_$1_$2.json 2> $O/b_$1_$2.err
  python - $O/b_$1_$2.json $1 $2 <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("in-flight %s stage %s: value %.1fM e2e %.1fM (in flight used %s)" % (sys.argv[2], sys.argv[3], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["setup"]["searches_in_flight"]))
except Exception as e:
    print("failed", sys.argv[1], e)
P
done
