#!/bin/bash
# quick regression + bench with hang guards: one Winograd unit first (a deadlocked kernel must not eat the GPU budget)
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd and conv1_1" > gpurun_out/quick_first.log 2>&1
rc=$?; echo "first exit $rc"; tail -3 gpurun_out/quick_first.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd or forward or infer_batch or conv_units or maxpool or upsample" > gpurun_out/quick_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/quick_tests.log
grep -E "passed|failed|exit|^FAILED" gpurun_out/quick_tests.log | tail -8
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1])
    print("value %.4g ms/step %.2f e2e %.4g launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]), d["clocks"])
    pu = d["roofline"]["per_unit"]
    print({k: round(v["ms_per_step"], 3) for k, v in pu.items()}, "conv total %.2f" % sum(v["ms_per_step"] for v in pu.values()))
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/quick_bench.err").read()[-2000:])
PY
