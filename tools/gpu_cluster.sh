#!/bin/bash
# first contact of the CTA-pair weight multicast (conv_wg_kernel<..., CL = 2>) with the hardware: hang-guarded unit test, the Winograd / forward /
# fused-call tests, bit-identity against SN_WG_CLUSTER=1, determinism, then an interleaved A/B of the two forms of the same build
mkdir -p gpurun_out
export SN_WG_CLUSTER=2      # the tests below run the opt-in form (the default is 1)
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "winograd and conv1_1" > gpurun_out/cl_first.log 2>&1
rc=$?; echo "first exit $rc"; tail -3 gpurun_out/cl_first.log
if [ $rc -ne 0 ]; then tail -30 gpurun_out/cl_first.log; exit 1; fi
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "winograd and (merge_conv2 or conv4_2 or conv1_2)" > gpurun_out/cl_second.log 2>&1
rc=$?; echo "second exit $rc"; tail -3 gpurun_out/cl_second.log
if [ $rc -ne 0 ]; then tail -30 gpurun_out/cl_second.log; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd or forward or infer_batch or fused or truncation" > gpurun_out/cl_tests.log 2>&1
echo "tests exit $?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/cl_tests.log | tail -12
timeout 300 python tools/determinism_check.py 32 6 8 2>&1 | tail -3
unset SN_WG_CLUSTER
for r in 1 2; do
for v in 2 1; do      # 2 = CTA pairs sharing a multicast weight stream, 1 = default
  SN_WG_CLUSTER=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/cl_ab_$v$r.json 2> gpurun_out/cl_ab_$v$r.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/cl_ab_$v$r.json").read().strip().splitlines()[-1])
    pu = d["roofline"]["per_unit"]
    print("cluster $v run $r ms/step %.2f e2e %.2f clock %s conv %.2f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"], sum(v["ms_per_step"] for v in pu.values())), {k: round(v["ms_per_step"], 2) for k, v in pu.items() if k in ("conv1_2","conv2_2","conv3_2","conv4_2","merge_conv","merge_conv2")})
except Exception as e:
    print("cluster $v unreadable", e); print(open("gpurun_out/cl_ab_$v$r.err").read()[-1500:])
PY
done; done
