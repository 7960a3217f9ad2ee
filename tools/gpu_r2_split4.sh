#!/bin/bash
# round 2 (4 B200s): the split exchange with more than one partner per rank: sharded parity with the defaults and the bench line
set -u
N=4
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DIST_CHECK_CONFIG3=0 timeout 500 $T --master-port 29604 tests/dist_check.py > gpurun_out/r2_dist_check_${N}gpu_split.log 2>&1
echo "dist_check rc=$?"; grep -c "err=" gpurun_out/r2_dist_check_${N}gpu_split.log; grep "DIST_CHECK\|FAIL\|config-5\|Error\|error" gpurun_out/r2_dist_check_${N}gpu_split.log | tail -6
timeout 300 $T --master-port 29624 bench.py --gpus $N --steps 4 --warmup 3 --e2e 0 > gpurun_out/r2_bench_${N}gpu_split24.json 2> gpurun_out/r2_bench_${N}gpu_split24.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_${N}gpu_split24.err
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_4gpu_split24.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["ms_per_step"],2), d["gpu_launches"], json.dumps(d["detail"]["parity_check"]), json.dumps(d["detail"]["nvlink_swaps"])[-420:])
except Exception as e: print("ERR", e)
P
