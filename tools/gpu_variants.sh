#!/bin/bash
# A/B of libqsv variants (tools/regs_variants.py) on one B200: parity of the fused path, then the config-2 circuit.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/variants.log
: > $OUT
B="python bench.py --steps 3 --warmup 3 --sweeps 0 --e2e 0 --cpu-baseline 0 --adjoint 0"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["ms_per_launch"], d["roofline"].get("hbm_actual_frac"))'
for v in default "$@"; do
  if [ "$v" = default ]; then unset QSV_LIB_PATH; else export QSV_LIB_PATH=$PWD/pennylane_lightning_gpu_b200/lib/variants/libqsv_$v.so; fi
  echo "== $v: parity" >> $OUT
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stress or random_circuit or large_state" 2>&1 | tail -2 >> $OUT
  echo "== $v: c128" >> $OUT
  timeout 300 $B 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
  echo "== $v: c64" >> $OUT
  timeout 300 $B --dtype c64 2>&1 | tail -1 | python -c "$P" >> $OUT 2>&1
done
cat $OUT
