#!/bin/bash
# round 2, call 2 (1 B200): full GPU test-suite on the persistent sweep kernel + the new batched generator kernel,
# A/B of both against their predecessors / variants, ncu launch lists and one full capture of each.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_call2.log
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 >> $OUT
B="python bench.py --steps 3 --warmup 3 --sweeps 0 --e2e 0 --cpu-baseline 0 --adjoint 0"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "GB/s", round(d["ms_per_step"],2), "ms/step", d["gpu_launches"]//d["steps"], "launches/step", round(d["roofline"]["ms_per_launch"],2), "ms/launch hbm_frac", round(d["roofline"].get("hbm_actual_frac") or 0,3), d["clocks"])'
run() {
  local label="$1"; shift
  echo "== $label" >> $OUT
  env "$@" timeout 300 $B 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
}
run "c128 persistent (default)" QSV_DUMMY=1
run "c128 generation 2 (QSV_REGS_PERSIST=0)" QSV_REGS_PERSIST=0
run "c128 persistent, 1 CTA/SM" QSV_REGS_PERSIST_CTAS=1
echo "== c64 persistent" >> $OUT
timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
echo "== c64 generation 2" >> $OUT
env QSV_REGS_PERSIST=0 timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
echo "== c64 persistent, 4 CTAs/SM" >> $OUT
env QSV_REGS_PERSIST_CTAS=4 timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
echo "== adjoint config 3" >> $OUT
A="python tools/ab_adjoint.py 24"
timeout 200 $A >> $OUT 2>&1
env QSV_GENS_TB=12 timeout 200 $A >> $OUT 2>&1
env QSV_GENS_L13=3 timeout 200 $A >> $OUT 2>&1
env QSV_GENS_L13=4 timeout 200 $A >> $OUT 2>&1
env QSV_ADJOINT_DEFER=0 timeout 200 $A >> $OUT 2>&1
env QSV_REGS_PERSIST=0 timeout 200 $A >> $OUT 2>&1
env QSV_ADJOINT_BATCH=0 timeout 200 $A >> $OUT 2>&1
# ncu: launch lists (shares), then one full capture per kernel family
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_adjoint.csv $A > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_fused.csv $B --steps 2 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_regs_p -s 9 -c 2 -o gpurun_out/r2_regs_p $B --steps 1 --warmup 1 > gpurun_out/r2_ncu_regs_p.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bra_gens_ket -s 4 -c 3 -o gpurun_out/r2_gens $A > gpurun_out/r2_ncu_gens.log 2>&1
cat $OUT
