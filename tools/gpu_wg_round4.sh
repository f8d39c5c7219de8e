#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd" -s > gpurun_out/wg_units.log 2>&1
echo "units exit $?" >> gpurun_out/wg_units.log
grep -E "winograd: max-abs|passed|failed|exit|Error|error" gpurun_out/wg_units.log | tail -30
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "forward or infer_batch" -s > gpurun_out/wg_forward.log 2>&1
echo "forward exit $?" >> gpurun_out/wg_forward.log
grep -E "max-abs|passed|failed|exit" gpurun_out/wg_forward.log | tail -20
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_wg_b.json 2> gpurun_out/r02_bench_wg_b.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_wg_b.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("parity_check"), d["clocks"])
        print({k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_unit"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/r02_bench_wg_b.err
OUT=gpurun_out/r02_survey_b.jsonl; rm -f $OUT
timeout 900 python tools/parity_survey.py --seeds 0 1 --nvp 2 --D 64 --cubes 1 --tag wg4 --out $OUT 2>&1 | grep -E "summary|Error|error"
