#!/bin/bash
# bring-up of the weight multicast: control (SN_WG_CLUSTER=1) and the other units under the cluster launch
mkdir -p gpurun_out
SN_WG_CLUSTER=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "winograd and (conv1_1 or merge_conv2-16)" > gpurun_out/clb_ctl.log 2>&1
echo "control exit $?"; grep -E "winograd: max-abs|passed|failed" gpurun_out/clb_ctl.log | tail -5
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "winograd" > gpurun_out/clb_all.log 2>&1
echo "cluster exit $?"; grep -E "winograd: max-abs|passed|failed" gpurun_out/clb_all.log | grep -v print | tail -24
