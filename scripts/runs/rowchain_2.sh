#!/bin/bash
# row-block executor v5: parity + stamps, the default bench line, and 512-tree shards with 10 / 20 searches in flight
O=gpurun_out/rc5; mkdir -p $O
timeout 240 python scripts/exp_rowchain.py > $O/exp.log 2>&1; echo "exp rc=$?"; grep -v "max|rows" $O/exp.log | tail -n 30
show() { python - "$1" "$2" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%s: value %.1fM e2e %.1fM one %.1fM in-flight %s executor %s host_us %.0f" % (sys.argv[2], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["one_search_at_a_time"]["value"] / 1e6, d["setup"]["searches_in_flight"], d["setup"].get("network_executor"), d["setup"]["host_us_per_submit"]))
except Exception as e:
    print("no bench line", sys.argv[1], e)
P
}
timeout 300 python bench.py --quick --no-cpu-baseline > $O/bench_4096.json 2> $O/bench_4096.err; echo "rc=$?"; tail -n 3 $O/bench_4096.err; show $O/bench_4096.json 4096
timeout 300 python bench.py --quick --no-cpu-baseline --trees 512 --in-flight 10 > $O/bench_512_10.json 2> $O/bench_512_10.err; echo "rc=$?"; tail -n 3 $O/bench_512_10.err; show $O/bench_512_10.json 512x10
timeout 300 python bench.py --quick --no-cpu-baseline --trees 512 > $O/bench_512_20.json 2> $O/bench_512_20.err; echo "rc=$?"; tail -n 3 $O/bench_512_20.err; show $O/bench_512_20.json 512x20
