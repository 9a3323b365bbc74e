#!/bin/bash
# the whole GPU-marked suite, the driver's smoke entry and the default bench line, with the row-block executor in
O=gpurun_out/final3; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 $O/pytest.log
timeout 600 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -n 3 $O/bench_default.err
python - <<'P'
import json
d = json.load(open("gpurun_out/final3/bench_default.json"))
print("value %.1fM e2e %.1fM one %.1fM" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["one_search_at_a_time"]["value"] / 1e6))
print("launches", d["gpu_launches_per_search"], d["library_gemm_launches_per_search"], d["launches_one_search_at_a_time"])
print("network_roofline", {k: v for k, v in (d["network_roofline"] or {}).items() if k in ("achieved", "frac", "launch_us_aggregate", "alone", "algorithmic_flops_per_row")})
print("plan_vs_module", d["plan_vs_module"]); print("selfplay", d["selfplay"]["simulations_per_sec"], d["selfplay"]["fraction_of_search_only"])
print("roofline frac", d["roofline"]["frac"], d["roofline"]["launch_us"])
P
