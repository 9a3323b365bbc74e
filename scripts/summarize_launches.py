"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for ONE whole search
(the launches between two k_root_stats).   python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/summary.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
data = [r for r in rows[hdr + 1:] if len(r) == len(H)]
ik, iv = H.index("Kernel Name"), H.index("Metric Value")
names = [r[ik] for r in data]
marks = [i for i, n in enumerate(names) if "k_root_stats" in n]
lo, hi = (marks[0] + 1, marks[1] + 1) if len(marks) >= 2 else (0, len(data))
agg = collections.OrderedDict()
for r in data[lo:hi]:
    name = re.sub(r"\(.*", "", r[ik])[:100]
    a = agg.setdefault(name, [0.0, 0])
    a[0] += float(r[iv]) / 1e3
    a[1] += 1
total = sum(a[0] for a in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold caches: compare SHARES)")
print(f"# one whole search = the launches between two k_root_stats; kernels in one search {hi - lo}  sum {total:.1f} us")
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:9.1f} us {100 * us / total:5.1f}%  n={n:4d} avg {us / n:6.2f}  {name}")
