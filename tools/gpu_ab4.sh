#!/bin/bash
# fourth A/B: staggered start of the CTAs that share an SM (QSV_REGS_STAGGER_NS: -1 = auto, 0 = off, > 0 = ns per slot)
set -u
mkdir -p gpurun_out
OUT=gpurun_out/ab4.log
: > $OUT
B="python bench.py --steps 3 --warmup 3 --sweeps 0 --e2e 0 --cpu-baseline 0 --adjoint 0"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "GB/s", round(d["ms_per_step"],2), "ms/step", d["gpu_launches"]//d["steps"], "launches/step", round(d["roofline"]["ms_per_launch"],2), "ms/launch hbm_frac", round(d["roofline"].get("hbm_actual_frac") or 0,3))'
run() {
  local label="$1"; shift
  echo "== $label" >> $OUT
  env "$@" timeout 300 $B 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
}
run "stagger off" QSV_REGS_STAGGER_NS=0
run "stagger auto" QSV_DUMMY=1
run "stagger 2000 ns" QSV_REGS_STAGGER_NS=2000
run "stagger 4000 ns" QSV_REGS_STAGGER_NS=4000
run "stagger 8000 ns" QSV_REGS_STAGGER_NS=8000
echo "== c64 off / auto / 3000" >> $OUT
for s in 0 -1 3000; do env QSV_REGS_STAGGER_NS=$s timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1; done
echo "== hea4 / sel2 30 qubits: off, auto, 2000" >> $OUT
for s in 0 -1 2000; do QSV_REGS_STAGGER_NS=$s python tools/bench_hea.py 30 2>&1 | tail -2 | cut -c1-200 >> $OUT; done
cat $OUT
