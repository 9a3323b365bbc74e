#!/bin/bash
# Evidence for the row-block executor: ncu --set full of k_row_chain, launch list of one whole search on it (each
# program exits 0 without ncu first), 2048-tree shard with 7 / 20 searches in flight, final bench lines of both arms.
O=gpurun_out/r4; mkdir -p $O
timeout 120 python scripts/prof_rowchain.py > $O/plain_rowchain.log 2>&1; echo "plain rowchain rc=$?"; cat $O/plain_rowchain.log | tail -1
timeout 300 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:k_row_chain --launch-skip 6 --launch-count 2 -o $O/r02_rowchain4096 python scripts/prof_rowchain.py > $O/ncu_rowchain.log 2>&1; echo "ncu rowchain rc=$?"
ncu -i $O/r02_rowchain4096.ncu-rep --page raw --csv > $O/r02_ncu_full_k_row_chain_4096.csv 2>/dev/null
timeout 300 python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --in-flight 1 --executor rows > $O/plain_bench_rows1.json 2> $O/plain_bench_rows1.err; echo "plain bench rows in-flight 1 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_ncu_launches_bench_rows.csv python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --in-flight 1 --executor rows > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for cfg in "2048 7" "2048 20"; do set -- $cfg
  timeout 300 python bench.py --quick --no-cpu-baseline --trees $1 --in-flight $2 > $O/bench_$1_$2.json 2> $O/bench_$1_$2.err; echo "bench $1 x $2 rc=$?"
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -n 3 $O/bench_default.err
rm -f $O/r02_rowchain4096.ncu-rep.tmp
python - <<'P'
import json
for f in ("bench_2048_7", "bench_2048_20", "bench_default"):
    try:
        d = json.load(open(f"gpurun_out/r4/{f}.json"))
        print(f, "value %.1fM e2e %.1fM one %.1fM in flight %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["one_search_at_a_time"]["value"] / 1e6, d["setup"]["searches_in_flight"]))
    except Exception as e:
        print(f, "ERR", e)
try:
    r = json.load(open("gpurun_out/r4/bench_reference.json")); d = json.load(open("gpurun_out/r4/bench_default.json"))
    print("reference %.2fM  e2e ratio %.1f" % (r["value"] / 1e6, d["e2e"]["value"] / r["value"]), d["plan_vs_module"].get("rows_executor"))
except Exception as e:
    print("ERR", e)
P
ls -la $O | head -30
