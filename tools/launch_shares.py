#!/usr/bin/env python
"""ncu launch list (csv from --metrics gpu__time_duration.sum) -> markdown share table of ONE step.
usage: tools/launch_shares.py <launches.csv> <title>"""
import collections, csv, re, sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(lines))
hdr = r[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "second": 1e3}
seq = [(re.sub(r"\(.*", "", x[ki]).replace("facl::<unnamed>::", ""), float(x[vi].replace(",", "")) * scale[x[ui]]) for x in r[1:]]
g = [i for i, (n, _) in enumerate(seq) if "group_kernel" in n]
step = seq[g[1]:g[2]] if len(g) > 2 else seq
tot = sum(v for _, v in step)
agg = collections.OrderedDict()
for n, v in step:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
print(f"# {sys.argv[2]}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` launch list of `python bench.py --steps 2 --warmup 1 "
      "--no-cpu-baseline` (cfg2); one training step (cold-cache, serialised: compare SHARES).\n")
print(f"step: {len(step)} launches, {tot:.3f} ms under ncu\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {v:.3f} | {100 * v / tot:.1f}% |")
