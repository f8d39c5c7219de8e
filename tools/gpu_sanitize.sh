#!/bin/bash
# compute-sanitizer over the hot-path kernels (conv_tc mbarrier/TMEM protocol, raypool atomics, cvc gather).  Run under gpurun;
# summaries land in gpurun_out/sanitize_*.log and are copied to profiles/ by hand.
set -u
export SN_TC_TUNE=0            # no autotune sweeps under the sanitizer (first candidate of every unit)
T="tests/test_gpu_parity.py"
SEL='test_conv_units and exact and (conv1_2 or conv4_1 or merge_conv2 or side_op1) or test_forward_single_pair_and_odd_batch and exact or test_raypool_matches_reference_outputs or test_cvc_matches_reference_outputs or test_cvc_index_map_bit_exact or test_infer_batch_host_matches_oracle_pipeline and exact'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $T -x -q -m gpu -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/sanitize_$tool.log | tail -5
done
