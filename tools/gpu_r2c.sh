#!/bin/bash
# round 2, second session, evidence run: bench (plain pass vs pass with per-launch events), sanitizers over the kernels this session
# changed (ray pooling, fused gather, side + pool), launch list, ncu --set full of the non-conv kernels of one step
mkdir -p gpurun_out
for r in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench$r.json 2> gpurun_out/r2c_bench$r.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c_bench$r.json").read().strip().splitlines()[-1])
    print("bench$r ms/step %.2f with events %.2f e2e %.2f sparse %.2f launches %d clock %s" % (d["ms_per_step"], d["roofline"]["ms_per_step_with_events"], d["e2e"]["ms_per_step"], d["e2e_sparse"]["ms_per_step"], d["gpu_launches"], d["clocks"]))
except Exception as e:
    print("bench$r unreadable", e); print(open("gpurun_out/r2c_bench$r.err").read()[-1500:])
PY
done
export SN_TC_TUNE=0
SEL='test_raypool_matches_reference_outputs or test_raypool_dense_selection or test_fused_gather_is_bit_identical and 16-2-1 or test_infer_batch_device_equals_host_entry or test_infer_batch_host_matches_oracle_pipeline and exact or test_cvc_matches_reference_outputs'
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > gpurun_out/r2c_sanitize_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/r2c_sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/r2c_sanitize_$tool.log | tail -4
done
unset SN_TC_TUNE
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r2c_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2c_launches.csv | tail -30
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"cvc_wino|side_pool|upsample_wino|pool_blk|rp_|raw_to_wino|fuse_kernel|cast_f32" -c 20 -o gpurun_out/r2c_small -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r2c_small.log 2>&1
tail -2 gpurun_out/r2c_small.log
ncu -i gpurun_out/r2c_small.ncu-rep --page raw --csv > gpurun_out/r2c_small_raw.csv 2>/dev/null; wc -c gpurun_out/r2c_small_raw.csv
ls -la gpurun_out/r2c_small.ncu-rep; rm -f gpurun_out/r2c_small.ncu-rep
