#!/bin/bash
# second A/B: two-qubit absorption, register-bit count, low tile bits; full GPU test-suite first
set -u
mkdir -p gpurun_out
OUT=gpurun_out/ab2.log
: > $OUT
B="python bench.py --steps 3 --warmup 3 --sweeps 0 --e2e 0 --cpu-baseline 0 --adjoint 0"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "GB/s", round(d["ms_per_step"],2), "ms/step", d["gpu_launches"]//d["steps"], "launches/step", round(d["roofline"]["ms_per_launch"],2), "ms/launch hbm_frac", round(d["roofline"].get("hbm_actual_frac") or 0,3))'
echo "== pytest -m gpu" >> $OUT
timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -E "^FAILED|^ERROR|passed|failed" | head -20 >> $OUT
run() {
  local label="$1"; shift
  echo "== $label" >> $OUT
  env "$@" timeout 300 $B 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
}
run "default (2q absorption on)" QSV_DUMMY=1
run "no 2q absorption" QSV_MERGE_2Q=0
run "RB=3" QSV_REGS_RB=3
run "L=3" QSV_REGS_LOW=3
run "L=5" QSV_REGS_LOW=5
run "no DAG" QSV_REGS_DAG=0
echo "== c64 default" >> $OUT
timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
echo "== c64 RB=3" >> $OUT
env QSV_REGS_RB=3 timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
echo "== c64 L=5" >> $OUT
env QSV_REGS_LOW=5 timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
cat $OUT
