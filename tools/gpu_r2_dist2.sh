#!/bin/bash
# round 2, call 1 (2 B200s): the sharded-register parity with the default (DAG) exchange schedule and with the fused
# exchange, then the weak-scaling bench line with and without QSV_DIST_FUSED_SWAP.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/r2_dist_pytest_2gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_dist_pytest_2gpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $T --master-port 29614 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu_default.json 2> gpurun_out/r2_bench_2gpu_default.err
echo "bench default rc=$?"; tail -c 300 gpurun_out/r2_bench_2gpu_default.err
QSV_DIST_FUSED_SWAP=1 timeout 400 $T --master-port 29615 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu_fused.json 2> gpurun_out/r2_bench_2gpu_fused.err
echo "bench fused rc=$?"; tail -c 300 gpurun_out/r2_bench_2gpu_fused.err
python - <<'P'
import json
for f in ("default","fused"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_2gpu_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], json.dumps(d.get("detail"))[:800])
    except Exception as e: print(f, "ERR", e)
P
