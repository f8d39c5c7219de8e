#!/bin/bash
# ncu --set full of every conv launch of ONE timed bench step (18 launches: 14 Winograd units, side_op1, the three direct 1x1x1 side units)
timeout 1200 ncu --set full --clock-control none --profile-from-start off -k regex:"conv_wg_kernel|conv_tc_kernel|side_wino_kernel" -c 18 -o gpurun_out/r02_ncu_all -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r02_ncu_all.log 2>&1
tail -2 gpurun_out/r02_ncu_all.log
ncu -i gpurun_out/r02_ncu_all.ncu-rep --page raw --csv > gpurun_out/r02_ncu_all_raw.csv 2>/dev/null; wc -c gpurun_out/r02_ncu_all_raw.csv
rm -f gpurun_out/r02_ncu_all.ncu-rep      # > 64 MiB: only the raw CSV travels back
