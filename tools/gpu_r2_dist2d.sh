#!/bin/bash
# round 2, call 12 (2 B200s): sharded parity (dist_check.py) and the weak-scaling line with interleaved staying / leaving
# tiles in the sweep that carries an exchange (and without, QSV_DIST_XCHG_INTERLEAVE=0)
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
DIST_CHECK_CONFIG3=0 timeout 600 $T --master-port 29604 tests/dist_check.py > gpurun_out/r2_dist_check_2gpu_b.log 2>&1
echo "dist_check rc=$?"; grep "DIST_CHECK\|FAIL\|config-5" gpurun_out/r2_dist_check_2gpu_b.log | tail -4
for mode in 1 0; do
  QSV_DIST_XCHG_INTERLEAVE=$mode timeout 400 $T --master-port 2961$mode bench.py --gpus 2 --steps 5 --warmup 3 --e2e 0 > gpurun_out/r2_bench_2gpu_il$mode.json 2> gpurun_out/r2_bench_2gpu_il$mode.err
  echo "bench interleave=$mode rc=$?"; tail -c 200 gpurun_out/r2_bench_2gpu_il$mode.err
done
python - <<'P'
import json
for f in ("il1","il0"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_2gpu_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), d["gpu_launches"], json.dumps(d["detail"]["parity_check"]), json.dumps(d["detail"]["nvlink_swaps"])[:400])
    except Exception as e: print(f, "ERR", e)
P
