#!/bin/bash
# end-of-round evidence run: full GPU test suite, smoke(), the bench lines that go to profiles/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -2 gpurun_out/r02_gpu_tests_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_exact_c3.json 2> gpurun_out/r02_bench_exact_c3.err; tail -c 700 gpurun_out/r02_bench_exact_c3.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_c3.json 2> /dev/null; tail -c 400 gpurun_out/r02_bench_reference_c3.json; echo
timeout 600 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r02_bench_exact_c2.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_exact_c2.json').read().strip().splitlines()[-1]); print('c2', d['value'], d['ms_per_step'])"
timeout 600 python bench.py --mode fast --no-cpu-baseline > gpurun_out/r02_bench_fast_c3.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_fast_c3.json').read().strip().splitlines()[-1]); print('fast', d['value'], d['ms_per_step'])"
timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 > gpurun_out/r02_bench_c5_n1.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c5_n1.json').read().strip().splitlines()[-1]); print('c5', d['value'], d['ms_per_step'])"
timeout 600 python bench.py --c4-shard --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4_shard_n1.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c4_shard_n1.json').read().strip().splitlines()[-1]); print('c4 shard', d['c4_shard'])"
