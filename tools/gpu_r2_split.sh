#!/bin/bash
# round 2 (N = 2 or 4 B200s; usage: gpu_r2_split.sh N): the split exchange (push a quarter in the sweep before, pull a quarter in
# the sweep after): sharded parity with it on (default), the weak-scaling line with it on and off
set -u
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DIST_CHECK_CONFIG3=0 timeout 600 $T --master-port 29604 tests/dist_check.py > gpurun_out/r2_dist_check_${N}gpu_split.log 2>&1
echo "dist_check rc=$?"; grep -c "err=" gpurun_out/r2_dist_check_${N}gpu_split.log; grep "DIST_CHECK\|FAIL\|config-5\|Error\|error" gpurun_out/r2_dist_check_${N}gpu_split.log | tail -6
DIST_CHECK_FUSED=1 timeout 300 $T --master-port 29605 tests/dist_check.py > gpurun_out/r2_dist_check_${N}gpu_split_fused.log 2>&1
echo "dist_check fused rc=$?"; grep "DIST_CHECK\|FAIL\|fused exchange" gpurun_out/r2_dist_check_${N}gpu_split_fused.log | tail -6
for mode in 24 0; do
  QSV_DIST_SPLIT_XCHG=$mode timeout 400 $T --master-port 2962$N bench.py --gpus $N --steps 4 --warmup 3 --e2e 0 > gpurun_out/r2_bench_${N}gpu_split$mode.json 2> gpurun_out/r2_bench_${N}gpu_split$mode.err
  echo "bench split=$mode rc=$?"; tail -c 300 gpurun_out/r2_bench_${N}gpu_split$mode.err
done
python - $N <<'P'
import json,sys
N=sys.argv[1]
for f in ("split24","split0"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_{N}gpu_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), d["gpu_launches"], json.dumps(d["detail"]["parity_check"]), json.dumps(d["detail"]["nvlink_swaps"])[-420:])
    except Exception as e: print(f, "ERR", e)
P
