#!/bin/bash
# first contact of conv_wg.cu with the hardware: unit tests, error maps, then the forward tests
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd" -s > gpurun_out/wg_units.log 2>&1
echo "units exit $?" >> gpurun_out/wg_units.log
grep -E "winograd: max-abs|passed|failed|exit|Error|error" gpurun_out/wg_units.log | tail -30
for u in "conv1_2 16" "merge_conv2 16" "conv2_2 32" "conv1_1 16"; do timeout 300 python tools/wg_debug.py $u 1 >> gpurun_out/wg_debug.log 2>&1; done
tail -60 gpurun_out/wg_debug.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "forward or infer_batch or conv_units" -s > gpurun_out/wg_forward.log 2>&1
echo "forward exit $?" >> gpurun_out/wg_forward.log
grep -E "max-abs|passed|failed|exit" gpurun_out/wg_forward.log | tail -40
