#!/bin/bash
# One B200: GPU parity tests, the default bench line, tile-low-bit sweeps, ncu launch list + one full capture.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/bench_full.log
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 >> $OUT
echo "== default bench" >> $OUT
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 600 gpurun_out/bench_default.err >> $OUT
python - >> $OUT <<'P'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
det=d.pop("detail",{})
print(json.dumps(d)[:1800])
for k,v in det.items():
    if k=="single_gate_sweeps":
        print(k, {g:(round(x["min_frac_of_peak"],3), round(x["median_frac_of_peak"],3)) for g,x in v.items()})
    else:
        print(k, json.dumps(v)[:700])
P
B="python bench.py --steps 3 --warmup 3 --sweeps 0 --e2e 0 --cpu-baseline 0 --adjoint 0"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["ms_per_launch"], d["roofline"].get("hbm_actual_frac"))'
for L in ${LS:-3}; do
  echo "== regs c128 L=$L" >> $OUT
  QSV_REGS_LOW=$L timeout 300 $B 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
done
for L in ${LS64:-4 5}; do
  echo "== regs c64 L=$L" >> $OUT
  QSV_REGS_LOW=$L timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
done
echo "== ncu launch list" >> $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches_fused.csv $B --steps 2 > /dev/null 2>&1
tail -2 gpurun_out/r1_launches_fused.csv | cut -c1-300 >> $OUT
echo "== ncu regs full" >> $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_regs -s 12 -c 2 -o gpurun_out/r1_regs -f $B --steps 1 > gpurun_out/ncu_regs.log 2>&1
tail -2 gpurun_out/ncu_regs.log | cut -c1-200 >> $OUT
cat $OUT
