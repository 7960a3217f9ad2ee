#!/bin/bash
# round 2, call 9 (1 B200): the default bench line as the driver runs it (timed), the reference arm, and one ncu --set full
# pass over every kernel family that is not the fused sweep (tools/profile_families.py)
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_call9.log
: > $OUT
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
echo "default bench rc=$? wall $(( $(date +%s) - T0 )) s" >> $OUT
python - >> $OUT 2>&1 <<'P'
import json
d=json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"])
print("roofline",{k:v for k,v in d["roofline"].items() if k in ("frac","frac_physical","frac_compute","hbm_actual_gbs")})
print("clocks",d["clocks"])
print("adjoint",d["detail"].get("adjoint_config3"))
print("config1",d["detail"].get("config1_sel20"))
print("config4",d["detail"].get("sparse_config4"))
print("state_io",d["detail"].get("state_io"))
print("c64",d["detail"].get("config2_complex64"))
print("unfused",d["detail"].get("unfused_gate_by_gate"))
P
T0=$(date +%s); timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
echo "reference arm rc=$? wall $(( $(date +%s) - T0 )) s: $(cut -c1-400 gpurun_out/r2_bench_reference.json)" >> $OUT
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2_families python tools/profile_families.py > gpurun_out/r2_families.log 2>&1
grep STEP gpurun_out/r2_families.log >> $OUT
cat $OUT
