#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the step).
    python profiles/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[im] != "gpu__time_duration.sum":
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    k = r[ik].split("(")[0][:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += float(r[iv].replace(",", "")) * scale
tot = sum(v[1] for v in agg.values())
lines = ["| kernel | launches | total ms (cold, serialised) | share |", "|---|---|---|---|"]
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append("| %s | %d | %.3f | %.1f %% |" % (k, n, ms, 100 * ms / tot))
lines.append("| total | %d | %.3f | |" % (sum(v[0] for v in agg.values()), tot))
txt = "\n".join(lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("# ncu launch list of bench.py's timed region (gpu__time_duration.sum, --clock-control none)\n\n" + txt + "\n")
