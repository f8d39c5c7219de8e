#!/bin/bash
# A/B bench (Winograd on/off) + parity survey under the round-toward-zero compensation settings
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_wg_a.json 2> gpurun_out/r02_bench_wg_a.err
SN_WG=0 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_nowg_a.json 2> gpurun_out/r02_bench_nowg_a.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_wg_a.json", "gpurun_out/r02_bench_nowg_a.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("parity_check"), d["clocks"])
        print({k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_unit"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
OUT=gpurun_out/r02_survey_a.jsonl; rm -f $OUT
python tools/parity_survey.py --seeds 0 1 --nvp 2 --D 64 --cubes 1 --tag wg --out $OUT 2>&1 | grep summary
SN_WG=0 python tools/parity_survey.py --seeds 0 1 --nvp 2 --D 64 --cubes 1 --tag nowg --out $OUT 2>&1 | grep summary
SN_WG_RZSCALE=0 python tools/parity_survey.py --seeds 0 1 --nvp 2 --D 64 --cubes 1 --tag wg_rz0 --out $OUT 2>&1 | grep summary
SN_WG_RZSCALE=0.5 python tools/parity_survey.py --seeds 0 1 --nvp 2 --D 64 --cubes 1 --tag wg_rz0.5 --out $OUT 2>&1 | grep summary
SN_WG_RZSCALE=1.5 python tools/parity_survey.py --seeds 0 1 --nvp 2 --D 64 --cubes 1 --tag wg_rz1.5 --out $OUT 2>&1 | grep summary
SN_WG_RZSCALE=2 python tools/parity_survey.py --seeds 0 1 --nvp 2 --D 64 --cubes 1 --tag wg_rz2 --out $OUT 2>&1 | grep summary
python tools/parity_survey.py --seeds 0 1 --nvp 2 --D 64 --cubes 1 --mode fp32 --tag fp32 --out $OUT 2>&1 | grep summary
