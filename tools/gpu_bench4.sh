#!/bin/bash
# 4-GPU bench (NCCL gather of float16 + votes) -- run with gpurun --gpus 4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n4.json").read().strip().splitlines()[-1])
    print("n4 value %.4g ms/step %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/r02_bench_n4.err").read()[-3000:])
PY
