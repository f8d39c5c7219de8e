#!/bin/bash
# 2-GPU bench (NCCL gather of float16 + votes, c4 shard) -- run with gpurun --gpus 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
    print("n2 value %.4g ms/step %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["clocks"]); print(d.get("c4_shard"))
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/r02_bench_n2.err").read()[-3000:])
PY
