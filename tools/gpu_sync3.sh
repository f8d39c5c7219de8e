#!/bin/bash
export SN_TC_TUNE=0
SN_WG_AD=2 timeout 300 compute-sanitizer --tool synccheck --print-limit 2 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_conv_units_winograd and conv1_2-32" > gpurun_out/sync3_ad2.log 2>&1
echo "== AD=2: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sync3_ad2.log | tr '\n' ' ')"; grep -E "Barrier error|at void|located" gpurun_out/sync3_ad2.log | head -3
timeout 300 compute-sanitizer --tool synccheck --print-limit 2 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_conv_units_winograd and (conv2_1-16 or merge_conv2-16 or conv4_1-16 or conv3_2-8)" > gpurun_out/sync3_others.log 2>&1
echo "== others: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sync3_others.log | tr '\n' ' ')"
SN_WG_UNITS=0xFFFFF8 timeout 300 compute-sanitizer --tool synccheck --print-limit 2 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_forward_s32_two_pairs and exact" > gpurun_out/sync3_fwd.log 2>&1
echo "== forward without conv1 wino: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sync3_fwd.log | tr '\n' ' ')"; grep -E "Barrier error|at void|located" gpurun_out/sync3_fwd.log | head -3
