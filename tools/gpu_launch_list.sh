#!/bin/bash
# ncu launch list (device time per launch, cold & serialised) of one timed bench step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r02_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches.csv")) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
seq = []
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    seq.append((r[ik][:60], ms))
print(len(seq), "launches; in order:")
for k, ms in seq[:60]: print("  %-60s %.3f" % (k, ms))
PY
