#!/bin/bash
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd and conv1_1" > gpurun_out/quick_first.log 2>&1
rc=$?; echo "first exit $rc"; tail -2 gpurun_out/quick_first.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "parity_distribution or truncation_extremes or import_swap" -s > gpurun_out/r02_newtests.log 2>&1
echo "newtests exit $?" >> gpurun_out/r02_newtests.log
grep -E "max-abs|passed|failed|exit|^FAILED" gpurun_out/r02_newtests.log | tail -50
