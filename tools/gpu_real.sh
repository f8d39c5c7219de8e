#!/bin/bash
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "winograd and conv1_1" > gpurun_out/quick_first.log 2>&1
rc=$?; echo "first exit $rc"; if [ $rc -ne 0 ]; then tail -5 gpurun_out/quick_first.log; exit 1; fi
timeout 1200 python -m pytest tests/test_gpu_real.py -q -m gpu -s > gpurun_out/r02_real.log 2>&1
echo "real exit $?" >> gpurun_out/r02_real.log
grep -E "scan9 real|dinoSparseRing:|cube [0-9]+:|passed|failed|exit|Error|^E " gpurun_out/r02_real.log | tail -30
