#!/bin/bash
# Is the dominant launch bound by L2 -> SM traffic?  Timing-only runs with the weight stream (SN_WG_DEBUG=4), the lo operand plane (8) or both (12)
# kept out of the L2 -> SM path (results are garbage; --no-cpu-baseline skips the parity check), interleaved with the normal build, per-unit times from the bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_select.py -q -m gpu -k "main_reconstruct_dropin" > gpurun_out/l2_test.log 2>&1; tail -1 gpurun_out/l2_test.log
for r in 1 2; do
for d in 0 4 8 12; do
  SN_WG_DEBUG=$d timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/l2_dbg${d}_$r.json 2> gpurun_out/l2_dbg${d}_$r.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/l2_dbg${d}_$r.json").read().strip().splitlines()[-1])
    pu = d["roofline"]["per_unit"]
    print("dbg $d run $r ms/step %.2f clock %s conv %.2f" % (d["ms_per_step"], d["clocks"]["sm_mhz"], sum(v["ms_per_step"] for v in pu.values())), {k: round(v["ms_per_step"], 2) for k, v in pu.items() if k in ("conv1_2","conv2_2","conv3_2","conv4_2","merge_conv","merge_conv2")})
except Exception as e:
    print("dbg $d unreadable", e); print(open("gpurun_out/l2_dbg${d}_$r.err").read()[-800:])
PY
done; done
