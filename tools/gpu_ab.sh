#!/bin/bash
# A/B of two library builds on the SAME box, interleaved: $1 = path of the other .so
for r in 1 2; do
for v in new old; do
  if [ $v = old ]; then export SN_LIB_PATH=$1; else unset SN_LIB_PATH; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v$r.json 2> gpurun_out/ab_$v$r.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$v$r.json").read().strip().splitlines()[-1])
    pu = d["roofline"]["per_unit"]
    print("$v$r ms/step %.2f clock %s conv %.2f" % (d["ms_per_step"], d["clocks"]["sm_mhz"], sum(v["ms_per_step"] for v in pu.values())), {k: round(v["ms_per_step"], 2) for k, v in pu.items() if k in ("conv1_1","conv1_2","conv1_3","conv2_2","conv4_2","merge_conv","merge_conv2")})
except Exception as e:
    print("$v$r unreadable", e)
PY
done; done
