#!/bin/bash
# round 2, call 11 (1 B200): GPU suite with the row-based probability kernel; measurement kernels timed; CSR lanes-per-row A/B
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_call11.log
: > $OUT
echo "== pytest -m gpu" >> $OUT
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -8 >> $OUT
echo "== measurement kernels" >> $OUT
timeout 300 python tools/ab_measure.py 30 >> $OUT 2>&1
env QSV_GENS_SMALL_TB=0 timeout 300 python tools/ab_measure.py 30 2>&1 | grep "Pauli words" >> $OUT
echo "== csr" >> $OUT
for L in auto 8 16; do
  if [ $L = auto ]; then timeout 300 python tools/ab_csr.py 22 2>&1 | tail -1 >> $OUT; else env QSV_CSR_LPR=$L timeout 300 python tools/ab_csr.py 22 2>&1 | tail -1 >> $OUT; fi
done
echo "== adjoint" >> $OUT
timeout 200 python tools/ab_adjoint.py 24 >> $OUT 2>&1
cat $OUT
