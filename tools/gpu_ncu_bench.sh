#!/bin/bash
# ncu --set full of the dominant launches AT BENCH SIZE (timed region of bench.py, one step): merge_conv2 (FINAL instance) and merge_conv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_wg_kernel -s 12 -c 2 -o gpurun_out/r02_ncu_merge -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r02_ncu_merge.log 2>&1
tail -2 gpurun_out/r02_ncu_merge.log
ncu -i gpurun_out/r02_ncu_merge.ncu-rep --page raw --csv > gpurun_out/r02_ncu_merge_raw.csv 2>/dev/null; wc -c gpurun_out/r02_ncu_merge_raw.csv
