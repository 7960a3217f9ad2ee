#!/bin/bash
# third A/B: 4x4 blocks on lane bits through FP64 tensor-core MMA (QSV_REGS_MMA)
set -u
mkdir -p gpurun_out
OUT=gpurun_out/ab3.log
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
run "default (MMA mode 2: merged 4x4 block per pass)" QSV_DUMMY=1
run "MMA off" QSV_REGS_MMA=0
run "MMA mode 1 (one 4x4 gate per pass)" QSV_REGS_MMA=1
run "MMA mode 2, L=5" QSV_REGS_LOW=5
echo "== adjoint config 3 + config 1" >> $OUT
timeout 600 python bench.py --steps 1 --warmup 3 --sweeps 0 --e2e 0 --cpu-baseline 0 --adjoint 1 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d["detail"]["adjoint_config3"])[:400]); print(json.dumps(d["detail"]["config1_sel20"])[:300])' >> $OUT 2>&1
echo "== ncu full, 2 launches (default)" >> $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_regs -s 12 -c 2 -o gpurun_out/r1d_regs -f $B --steps 1 > gpurun_out/ncu_regs.log 2>&1
tail -2 gpurun_out/ncu_regs.log | cut -c1-200 >> $OUT
cat $OUT
