#!/bin/bash
# A/B of the fused executors on one B200: parity tests, then the config-2 circuit with the register-blocked
# kernel at several tile-low-bit settings against the first-generation TMA tile kernel, then one ncu capture.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/ab_regs.log
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 >> $OUT
B="python bench.py --steps 3 --warmup 3 --sweeps 0 --e2e 0 --cpu-baseline 0 --adjoint 0"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["ms_per_launch"], d["roofline"].get("hbm_actual_frac"))'
for L in ${LS:-4 5 6 7}; do
  echo "== regs c128 L=$L" >> $OUT
  QSV_REGS_LOW=$L timeout 300 $B 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
done
if [ "${TMA:-1}" = 1 ]; then
echo "== tma-tile c128" >> $OUT
QSV_TILE_KERNEL=0 timeout 300 $B 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
fi
for L in ${LS64:-5 6 7}; do
  echo "== regs c64 L=$L" >> $OUT
  QSV_REGS_LOW=$L timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
done
if [ "${TMA:-1}" = 1 ]; then
echo "== tma-tile c64" >> $OUT
QSV_TILE_KERNEL=0 timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
fi
echo "== ncu regs" >> $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_regs -s 20 -c 2 -o gpurun_out/r1_regs -f $B --steps 1 > gpurun_out/ncu_regs.log 2>&1
tail -3 gpurun_out/ncu_regs.log | cut -c1-300 >> $OUT
cat $OUT
