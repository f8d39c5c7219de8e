#!/bin/bash
# 8-GPU bench: c3 weak scaling line (+ c4 shard = BASELINE configs[3]) and the whole-scene c5 (BASELINE configs[4]); run with gpurun --gpus 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
    print("n8 value %.4g ms/step %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["clocks"]); print(d.get("c4_shard"))
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/r02_bench_n8.err").read()[-3000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 2 --warmup 1 --workload c5 > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err
tail -c 1500 gpurun_out/r02_bench_c5_n8.json; tail -3 gpurun_out/r02_bench_c5_n8.err
