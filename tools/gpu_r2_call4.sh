#!/bin/bash
# round 2, call 4 (1 B200): GPU suite (tightened tolerances, staged host copies, large-size parity) with error margins
# logged; batched generator kernel v3 A/B; pageable vs pinned host copy rates; ncu of the generator kernel
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_call4.log
: > $OUT
rm -f gpurun_out/r2_parity_margins.txt
echo "== pytest -m gpu" >> $OUT
QSV_TEST_MARGINS=gpurun_out/r2_parity_margins.txt timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -25 >> $OUT
echo "== adjoint config 3" >> $OUT
A="python tools/ab_adjoint.py 24"
timeout 200 $A >> $OUT 2>&1
env QSV_GENS_TB=12 timeout 200 $A >> $OUT 2>&1
env QSV_GENS_L13=3 timeout 200 $A >> $OUT 2>&1
env QSV_ADJOINT_DEFER=0 timeout 200 $A >> $OUT 2>&1
echo "== host copies" >> $OUT
timeout 300 python tools/ab_state_io.py >> $OUT 2>&1
env QSV_IO_THREADS=4 timeout 300 python tools/ab_state_io.py >> $OUT 2>&1
env QSV_IO_THREADS=16 timeout 300 python tools/ab_state_io.py >> $OUT 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bra_gens_ket -s 10 -c 2 -o gpurun_out/r2_gens_v3 $A > gpurun_out/r2_ncu_gens_v3.log 2>&1
cat $OUT
