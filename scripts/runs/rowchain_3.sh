#!/bin/bash
# more hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS, default 8) for more than 8 searches in flight
O=gpurun_out/rc6; mkdir -p $O
show() { python - "$1" "$2" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%s: value %.1fM e2e %.1fM one %.1fM in-flight %s executor %s host_us %.0f" % (sys.argv[2], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["one_search_at_a_time"]["value"] / 1e6, d["setup"]["searches_in_flight"], d["setup"].get("network_executor"), d["setup"]["host_us_per_submit"]))
except Exception as e:
    print("no bench line", sys.argv[1], e)
P
}
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python bench.py --quick --no-cpu-baseline --trees 512 > $O/bench_512_20.json 2> $O/bench_512_20.err; echo "rc=$?"; tail -n 3 $O/bench_512_20.err; show $O/bench_512_20.json 512x20_conn32
timeout 300 python bench.py --quick --no-cpu-baseline --trees 512 --executor library > $O/bench_512_20_lib.json 2> $O/bench_512_20_lib.err; echo "rc=$?"; tail -n 3 $O/bench_512_20_lib.err; show $O/bench_512_20_lib.json 512x20_conn32_library
timeout 300 python bench.py --quick --no-cpu-baseline --trees 1024 > $O/bench_1024_20.json 2> $O/bench_1024_20.err; echo "rc=$?"; show $O/bench_1024_20.json 1024x20_conn32
timeout 300 python bench.py --quick --no-cpu-baseline > $O/bench_4096.json 2> $O/bench_4096.err; echo "rc=$?"; show $O/bench_4096.json 4096_conn32
