#!/bin/bash
# One B200: A/B of the single-target dense kernel's launch shape, then the default bench line with the best shape.
set -u
mkdir -p gpurun_out
timeout 110 python tools/ab_dense1.py > gpurun_out/ab_dense1_best.txt 2> gpurun_out/ab_dense1.log
echo "ab rc=$?"; grep "^shape" gpurun_out/ab_dense1.log | cut -c1-260
BEST=$(tail -1 gpurun_out/ab_dense1_best.txt)
case "$BEST" in ''|*[!0-9]*) BEST=0;; esac
echo "best shape: $BEST"
export QSV_DENSE1_SHAPE=$BEST
timeout 170 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"; tail -c 500 gpurun_out/bench_final.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
det=d.pop("detail",{})
print(d["value"], d["ms_per_step"], d["e2e"], d["cpu_baseline"]["value"])
for k,v in det.items():
    if k=="single_gate_sweeps":
        print(k, {g:(round(x["min_frac_of_peak"],3), round(x["median_frac_of_peak"],3)) for g,x in v.items()})
    else:
        print(k, json.dumps(v)[:400])
P
