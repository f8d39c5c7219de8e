#!/bin/bash
# end-of-round evidence run: full GPU test suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -2 gpurun_out/r02_gpu_tests_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_exact_c3.json 2> gpurun_out/r02_bench_exact_c3.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_exact_c3.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['parity_check'], d['clocks'])"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_c3.json 2> /dev/null; tail -c 300 gpurun_out/r02_bench_reference_c3.json; echo
