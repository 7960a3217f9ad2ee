#!/bin/bash
# round 2 (N = 4 or 8 B200s; usage: gpu_r2_multi.sh N): sharded parity with the defaults (dist_check.py: DAG exchange
# schedule, fused exchanges, rank-0-only read, config-5 generator at 28 qubits vs 1 GPU vs CPU port, layered sharded adjoint),
# the weak-scaling line at 30 local qubits, and BASELINE config 5 (33 local qubits: 36 / 35 qubits on 8 / 4 GPUs)
set -u
N=${1:-8}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $T --master-port 29604 tests/dist_check.py > gpurun_out/r2_dist_check_${N}gpu.log 2>&1
echo "dist_check rc=$?"; grep -c "err=" gpurun_out/r2_dist_check_${N}gpu.log; grep "DIST_CHECK\|FAIL\|config-5\|config-3\|rank-0-only" gpurun_out/r2_dist_check_${N}gpu.log | tail -6
timeout 400 $T --master-port 29614 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_${N}gpu.err
timeout 900 $T --master-port 29624 bench.py --gpus $N --qubits 33 --steps 2 --warmup 3 > gpurun_out/r2_bench_${N}gpu_config5.json 2> gpurun_out/r2_bench_${N}gpu_config5.err
echo "config5 rc=$?"; tail -c 300 gpurun_out/r2_bench_${N}gpu_config5.err
python - $N <<'P'
import json,sys
N=sys.argv[1]
for f in (f"r2_bench_{N}gpu",f"r2_bench_{N}gpu_config5"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), "GB/s", round(d["ms_per_step"],1), "ms/step", d["config"]["workload"], "launches", d["gpu_launches"], json.dumps(d.get("detail"))[:1800])
    except Exception as e: print(f, "ERR", e)
P
