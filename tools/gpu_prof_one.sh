#!/bin/bash
# ncu --set full + source on ONE launch of the Winograd kernel: $1 = launch index inside a forward (0 conv1_1 .. 12 merge_conv, 13 merge_conv2), $2 = tag
SKIP=$((14 + $1))
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wg_kernel -s $SKIP -c 1 -o gpurun_out/r02_wg_one_$2 -f python tools/wg_profile.py 8 64 > gpurun_out/r02_wg_one_$2.log 2>&1
tail -2 gpurun_out/r02_wg_one_$2.log
