#!/bin/bash
export SN_TC_TUNE=0
timeout 300 compute-sanitizer --tool synccheck --print-limit 2 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_conv_units_winograd and (conv2_1-16 or conv1_2-32)" > gpurun_out/sync4.log 2>&1
echo "== $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sync4.log | tr '\n' ' ')"; grep -E "Barrier error|at void|located" gpurun_out/sync4.log | head -3
