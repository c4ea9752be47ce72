"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (last step only).
    python profiles/summarize_launches.py gpurun_out/train_launches.csv [--list]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:]]
start = [i for i, (k, _) in enumerate(data) if "patchify" in k][-1]
step = data[start:]
clean = lambda k: re.sub(r"\(.*", "", k).replace("void ", "").replace("vb::<unnamed>::", "")
agg = collections.OrderedDict()
for k, v in step:
    agg.setdefault(clean(k), [0.0, 0])
    agg[clean(k)][0] += v
    agg[clean(k)][1] += 1
tot = sum(v for v, _ in agg.values())
for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v / 1e3:9.1f} us  x{c:3d}  {100 * v / tot:5.1f}%  {k[:90]}")
print(f"{tot / 1e3:9.1f} us total, {len(step)} launches")
if "--list" in sys.argv:
    for k, v in step:
        print(f"{v / 1e3:8.1f}  {clean(k)[:80]}")
