#!/usr/bin/env python
"""Table of the HBM-bound helper kernels from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` log: duration, DRAM bytes, achieved GB/s vs the measured copy peak (MEASURED_PEAKS.json hbm_gbs).
    python profiles/summarize_small.py <log.csv> [out.md]"""
import csv, json, os, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu, im, iid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name"), hdr.index("ID")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}
launches = collections.OrderedDict()
for r in rows:
    if r is hdr:
        continue
    d = launches.setdefault(r[iid], {"name": r[ik].split("(")[0][:60]})
    d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
peak = 6556.2
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
lines = ["| kernel | ms (ncu, cold) | DRAM read MB | DRAM write MB | achieved GB/s | of measured %.0f GB/s |" % peak, "|---|---|---|---|---|---|"]
for d in launches.values():
    t, rd, wr = d.get("gpu__time_duration.sum", 0), d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
    gbs = (rd + wr) / t / 1e9 if t else 0
    lines.append("| %s | %.3f | %.1f | %.1f | %.0f | %.1f %% |" % (d["name"], t * 1e3, rd / 1e6, wr / 1e6, gbs, 100 * gbs / peak))
txt = "\n".join(lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("# HBM-bound helper kernels of one bench.py step (ncu, --clock-control none; cold caches)\n\n" + txt + "\n")
