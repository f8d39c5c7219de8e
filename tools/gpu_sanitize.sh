#!/bin/bash
# compute-sanitizer over the hot-path kernels: conv_wg / conv_tc (mbarrier + TMEM protocols, setmaxnreg), raypool (atomics), cvc gather,
# the fused side / up-sample passes.  Run under gpurun; summaries land in gpurun_out/sanitize_*.log and are copied to profiles/.
set -u
export SN_TC_TUNE=0            # no autotune sweeps under the sanitizer (first candidate of every direct unit)
T="tests/test_gpu_parity.py"
SEL='test_conv_units_winograd and (conv1_2-32 or conv2_1 or conv4_1-16 or conv3_2-8 or merge_conv2-16) or test_conv_units and exact and (side_op1 or conv4_1 or merge_conv2 or conv1_1) or test_forward_single_pair_and_odd_batch and exact or test_forward_s32_two_pairs and exact or test_raypool_matches_reference_outputs or test_cvc_matches_reference_outputs or test_cvc_index_map_bit_exact or test_infer_batch_host_matches_oracle_pipeline and exact'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $T -x -q -m gpu -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/sanitize_$tool.log | tail -4
done
