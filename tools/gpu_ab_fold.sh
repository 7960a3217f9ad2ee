#!/bin/bash
# A/B of the fused executor's features on one B200.  Host-side features are selectable by environment (one build):
# DAG sweep packing, folded index permutations, merged thread-uniform diagonal gates, diagonal-as-2x2, L2 prefetch,
# unpredicated dense paths (QSV_REGS_UCONST); lib/variants/libqsv_plainsmem.so = the same with -DQSV_PLAIN_SMEM.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/ab_fold.log
: > $OUT
B="python bench.py --steps 3 --warmup 3 --sweeps 0 --e2e 0 --cpu-baseline 0 --adjoint 0"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "GB/s", round(d["ms_per_step"],2), "ms/step", d["gpu_launches"]//d["steps"], "launches/step", round(d["roofline"]["ms_per_launch"],2), "ms/launch hbm_frac", round(d["roofline"].get("hbm_actual_frac") or 0,3))'
VAR=$PWD/pennylane_lightning_gpu_b200/lib/variants/libqsv_plainsmem.so
echo "== pytest -m gpu" >> $OUT
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 >> $OUT
run() {  # label, env assignments...
  local label="$1"; shift
  echo "== $label" >> $OUT
  env "$@" timeout 300 $B 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
}
run "old mode (program order, no fold, no diag merge, no diag1, no plain path)" QSV_REGS_DAG=0 QSV_REGS_FOLD=0 QSV_REGS_UDIAG=0 QSV_REGS_DIAG1=0 QSV_REGS_UCONST=0
run "all host features, predicated dense paths only" QSV_REGS_UCONST=0
run "all (default: plain dense paths, constants from the parameter bank)" QSV_DUMMY=1
run "all, plain dense paths with constants from shared memory" QSV_LIB_PATH=$VAR
run "all, L=3" QSV_REGS_LOW=3
run "all, L=3, plainsmem" QSV_REGS_LOW=3 QSV_LIB_PATH=$VAR
run "all, no fold" QSV_REGS_FOLD=0
echo "== c64 old mode" >> $OUT
env QSV_REGS_DAG=0 QSV_REGS_FOLD=0 QSV_REGS_UDIAG=0 QSV_REGS_DIAG1=0 QSV_REGS_UCONST=0 timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
echo "== c64 all, predicated only" >> $OUT
env QSV_REGS_UCONST=0 timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
echo "== c64 all (default)" >> $OUT
timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
echo "== c64 all plainsmem" >> $OUT
env QSV_LIB_PATH=$VAR timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
echo "== adjoint config 3 + config 1 (default features)" >> $OUT
timeout 600 python bench.py --steps 1 --warmup 3 --sweeps 0 --e2e 0 --cpu-baseline 0 --adjoint 1 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d["detail"]["adjoint_config3"])); print(json.dumps(d["detail"]["config1_sel20"]))' >> $OUT 2>&1
echo "== ncu full, 2 launches of k_tile_regs (default features)" >> $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_regs -s 12 -c 2 -o gpurun_out/r1c_regs -f $B --steps 1 > gpurun_out/ncu_regs.log 2>&1
tail -2 gpurun_out/ncu_regs.log | cut -c1-300 >> $OUT
cat $OUT
