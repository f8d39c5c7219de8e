#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd or forward or infer_batch" -s > gpurun_out/wg_tests5.log 2>&1
echo "tests exit $?" >> gpurun_out/wg_tests5.log
grep -E "passed|failed|exit|^FAILED" gpurun_out/wg_tests5.log | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_wg_c.json 2> gpurun_out/r02_bench_wg_c.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_wg_c.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["clocks"])
        print({k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_unit"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
ncu --set full --clock-control none --import-source on -k regex:conv_wg_kernel -s 14 -c 14 -o gpurun_out/r02_wg_prof2 -f python tools/wg_profile.py 8 64 > gpurun_out/r02_wg_prof2.log 2>&1
tail -2 gpurun_out/r02_wg_prof2.log
