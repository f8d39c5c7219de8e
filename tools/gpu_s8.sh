#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd" -s > gpurun_out/wg_units.log 2>&1
rc=$?; echo "units exit $rc"; grep -E "winograd: max-abs|passed|failed" gpurun_out/wg_units.log | tail -22
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "forward or infer_batch or conv_units or parity_distribution_s32 or truncation or import_swap or reconstruct" -s > gpurun_out/quick_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/quick_tests.log
grep -E "s=32 seed|mode exact|passed|failed|exit|^FAILED" gpurun_out/quick_tests.log | tail -20
