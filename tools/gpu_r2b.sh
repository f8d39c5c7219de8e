#!/bin/bash
# round 2, second session: fused gather / side+pool / ray-pool table sizing / D2H side stream -- targeted tests with hang guards, then an
# interleaved A/B of the new build against the previous one (tools/ab/lib_r2a.so) on the same box.  The other library is not tracked (*.so); rebuild it with
#   mkdir -p gpurun_out/oldbuild && git archive <rev> surfacenet_b200/csrc include | tar -x -C gpurun_out/oldbuild &&
#   make -C gpurun_out/oldbuild/surfacenet_b200/csrc && mkdir -p tools/ab && cp gpurun_out/oldbuild/surfacenet_b200/libsurfacenet_b200.so tools/ab/lib_r2a.so
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "raypool" > gpurun_out/r2b_first.log 2>&1
rc=$?; echo "raypool exit $rc"; tail -3 gpurun_out/r2b_first.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/r2b_first.log; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused_gather or fused_passes or infer_batch or cvc or dense2sparse or forward_s32 or forward_s64 or maxpool" > gpurun_out/r2b_tests.log 2>&1
echo "tests exit $?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/r2b_tests.log | tail -12
for r in 1 2; do
for v in new old; do
  if [ $v = old ]; then export SN_LIB_PATH=$PWD/tools/ab/lib_r2a.so; else unset SN_LIB_PATH; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ab_$v$r.json 2> gpurun_out/r2b_ab_$v$r.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2b_ab_$v$r.json").read().strip().splitlines()[-1])
    pu = d["roofline"]["per_unit"]
    print("$v$r ms/step %.2f e2e %.2f sparse %.2f launches %d clock %s conv %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_sparse"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"], sum(v["ms_per_step"] for v in pu.values())), {k: round(v["ms_per_step"], 2) for k, v in pu.items() if k in ("conv1_1","side_op1","conv2_1","merge_conv","merge_conv2")})
except Exception as e:
    print("$v$r unreadable", e); print(open("gpurun_out/r2b_ab_$v$r.err").read()[-1500:])
PY
done; done
