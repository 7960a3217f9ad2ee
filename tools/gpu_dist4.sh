#!/bin/bash
# 4 B200s: sharded-register parity (tests/dist_check.py) and the weak-scaling bench line at 4 GPUs.
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $T --master-port 29604 tests/dist_check.py > gpurun_out/dist_check_4gpu.log 2>&1
echo "dist_check rc=$?"; grep -c "err=" gpurun_out/dist_check_4gpu.log; grep "DIST_CHECK\|FAIL" gpurun_out/dist_check_4gpu.log | tail -5
timeout 300 $T --master-port 29614 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
echo "bench rc=$?"; tail -c 400 gpurun_out/bench_4gpu.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_4gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["gpu_launches"], json.dumps(d.get("detail"))[:600])
P
