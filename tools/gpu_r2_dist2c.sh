#!/bin/bash
# round 2, call 6 (2 B200s): sharded parity incl. the layered sharded adjoint (config 3 at 26 qubits), bench line at 2 GPUs,
# and the >= 32-local-qubit path (borrowed buffer, in-place exchanges) with the config-5 circuit at 33 qubits
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/r2_dist_pytest_2gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_dist_pytest_2gpu.log
grep -h "config-5\|config-3\|rank-0-only\|DIST_CHECK\|FAIL" gpurun_out/dist_check_2gpu.log gpurun_out/dist_check_fused_2gpu.log | tail -12
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $T --master-port 29614 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_2gpu.err
timeout 600 $T --master-port 29615 bench.py --gpus 2 --qubits 32 --steps 2 --warmup 3 > gpurun_out/r2_bench_2gpu_33q_config5.json 2> gpurun_out/r2_bench_2gpu_33q_config5.err
echo "bench 33q rc=$?"; tail -c 300 gpurun_out/r2_bench_2gpu_33q_config5.err
python - <<'P'
import json
for f in ("r2_bench_2gpu","r2_bench_2gpu_33q_config5"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["config"]["workload"], d["e2e"], json.dumps(d.get("detail"))[:1500])
    except Exception as e: print(f, "ERR", e)
P
