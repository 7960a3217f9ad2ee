#!/bin/bash
# round 2, call 7 (1 B200): GPU suite, adjoint with the tile-based Pauli-sum apply and the sign-flip generator path, the
# default bench line as the driver runs it (timed), A/B of the in-place 4x4 variant
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_call7.log
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -8 >> $OUT
echo "== adjoint config 3" >> $OUT
A="python tools/ab_adjoint.py 24"
timeout 200 $A >> $OUT 2>&1
env QSV_PAULI_SUM_TILED=0 timeout 200 $A >> $OUT 2>&1
env QSV_GENS_TB=12 timeout 200 $A >> $OUT 2>&1
echo "== default bench (as the driver runs it)" >> $OUT
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
echo "rc=$? wall $(( $(date +%s) - T0 )) s" >> $OUT
python - >> $OUT 2>&1 <<'P'
import json
d=json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"])
print("roofline",{k:v for k,v in d["roofline"].items() if k in ("frac","frac_physical","frac_compute","hbm_actual_gbs")})
print("clocks",d["clocks"])
print("adjoint",d["detail"].get("adjoint_config3"))
print("config4",d["detail"].get("sparse_config4"))
print("state_io",d["detail"].get("state_io"))
print("c64",d["detail"].get("config2_complex64"))
P
B="python bench.py --steps 3 --warmup 3 --sweeps 0 --e2e 0 --cpu-baseline 0 --adjoint 0 --config4 0"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "GB/s", round(d["ms_per_step"],2), "ms/step", d["gpu_launches"]//d["steps"], "launches/step", d["clocks"]["sm_mhz"])'
echo "== A/B in-place 4x4 (variant library)" >> $OUT
timeout 300 $B 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
env QSV_LIB_PATH=pennylane_lightning_gpu_b200/lib/variants/libqsv_d2inplace.so timeout 300 $B 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_adjoint.csv $A > /dev/null 2>&1
cat $OUT
