#!/bin/bash
# Two B200s: the bench line at 2 GPUs including the end-to-end leg (every rank uploads its shard from pinned memory).
set -u
mkdir -p gpurun_out
timeout ${T:-70} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_e2e.json 2> gpurun_out/bench_2gpu_e2e.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_2gpu_e2e.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_2gpu_e2e.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], json.dumps(d.get("detail"))[:500])
P
