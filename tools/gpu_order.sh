#!/bin/bash
for r in 1 2; do for o in 0 1; do
SN_WG_ORDER=$o timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/order_$o$r.json 2> gpurun_out/order_$o$r.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/order_$o$r.json").read().strip().splitlines()[-1])
    pu = d["roofline"]["per_unit"]
    print("order=$o run $r ms/step %.2f clock %s conv %.2f" % (d["ms_per_step"], d["clocks"]["sm_mhz"], sum(v["ms_per_step"] for v in pu.values())), {k: round(v["ms_per_step"], 2) for k, v in pu.items() if k in ("conv1_2","conv2_2","conv4_2","merge_conv","merge_conv2")})
except Exception as e:
    print("unreadable", e)
PY
done; done
SN_WG_ORDER=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd or forward_s32" 2>&1 | tail -2
