#!/bin/bash
export SN_TC_TUNE=0
for k in "conv1_2-32" "conv1_3-16" "conv1_2-64" "conv1_2-32"; do
  timeout 300 compute-sanitizer --tool synccheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_conv_units_winograd and $k" > gpurun_out/sync2_$k.log 2>&1
  echo "== $k: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sync2_$k.log | tr '\n' ' ')"
  grep -E "Barrier error|by thread|conv_wg.cu" gpurun_out/sync2_$k.log | head -4
done
