#!/bin/bash
for d in 0 1 3; do
SN_WG_DEBUG=$d timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/dbg_bench_$d.json 2> gpurun_out/dbg_bench_$d.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/dbg_bench_$d.json").read().strip().splitlines()[-1])
    pu = d["roofline"]["per_unit"]
    print("dbg=$d ms/step %.2f clock %s" % (d["ms_per_step"], d["clocks"]["sm_mhz"]), {k: round(v["ms_per_step"], 2) for k, v in pu.items() if k in ("conv1_1","conv1_2","conv1_3","conv2_2","conv4_2","merge_conv","merge_conv2")})
except Exception as e:
    print("unreadable", e)
PY
done
