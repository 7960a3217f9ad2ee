#!/bin/bash
# round 2, call 3 (2 B200s): sharded parity with the new defaults (fused exchange routed per element, lazy timing, map
# policy, rank-0-only read, config-5 generator at 28 qubits vs 1 GPU vs CPU port) and the weak-scaling line at 2 GPUs
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/r2_dist_pytest_2gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_dist_pytest_2gpu.log
grep -h "config-5\|rank-0-only\|DIST_CHECK\|fused exchange" gpurun_out/dist_check_2gpu.log gpurun_out/dist_check_fused_2gpu.log | tail -12
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in fused inplace; do
  if [ $mode = fused ]; then export QSV_DIST_FUSED_SWAP=1; else export QSV_DIST_FUSED_SWAP=0; fi
  timeout 400 $T --master-port 29614 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu_$mode.json 2> gpurun_out/r2_bench_2gpu_$mode.err
  echo "bench $mode rc=$?"; tail -c 300 gpurun_out/r2_bench_2gpu_$mode.err
done
python - <<'P'
import json
for f in ("fused","inplace"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_2gpu_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"], json.dumps(d.get("detail"))[:1200])
    except Exception as e: print(f, "ERR", e)
P
