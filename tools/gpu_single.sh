#!/bin/bash
timeout 60 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd and conv1_1" > gpurun_out/quick_first.log 2>&1
rc=$?; echo "first exit $rc"; if [ $rc -ne 0 ]; then tail -5 gpurun_out/quick_first.log; exit 1; fi
timeout 90 python tools/determinism_check.py 32 6 20 2>&1 | tail -2
timeout 90 python tools/determinism_check.py 64 4 8 2>&1 | tail -1
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd or device_equals_host or sparse_matches" 2>&1 | tail -2
for v in new old new old; do
  if [ $v = old ]; then export SN_LIB_PATH=$1; else unset SN_LIB_PATH; fi
  timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
    pu = d["roofline"]["per_unit"]
    print("$v ms/step %.2f clock %s conv %.2f" % (d["ms_per_step"], d["clocks"]["sm_mhz"], sum(v["ms_per_step"] for v in pu.values())), {k: round(v["ms_per_step"], 2) for k, v in pu.items() if k in ("conv1_1","conv1_2","conv2_2","conv4_2","merge_conv","merge_conv2")})
except Exception as e:
    print("$v unreadable", e)
PY
done
