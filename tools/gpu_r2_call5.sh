#!/bin/bash
# round 2, call 5 (1 B200): GPU suite with NO tolerance widening, generator kernel v3b, device mirror with fused apply,
# launch list of the adjoint
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_call5.log
: > $OUT
rm -f gpurun_out/r2_parity_margins.txt
echo "== pytest -m gpu" >> $OUT
QSV_TEST_MARGINS=gpurun_out/r2_parity_margins.txt timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -25 >> $OUT
echo "== adjoint config 3" >> $OUT
A="python tools/ab_adjoint.py 24"
timeout 200 $A >> $OUT 2>&1
env QSV_GENS_TB=12 timeout 200 $A >> $OUT 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_adjoint.csv $A > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bra_gens_ket -s 10 -c 2 -o gpurun_out/r2_gens_v3b $A > gpurun_out/r2_ncu_gens_v3b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pauli_sum_apply -s 2 -c 1 -o gpurun_out/r2_pauli_sum $A > gpurun_out/r2_ncu_pauli_sum.log 2>&1
cat $OUT
