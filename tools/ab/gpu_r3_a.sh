#!/bin/bash
# round 2, call A: full GPU test-suite on the new code, then TS vs QS (streamed / half / fully resident
# queries) at the 8-GPU shard size and at the full CAsT size, isotropic and common-mean data.
mkdir -p gpurun_out
nvidia-smi -L | head -2
echo "=== gpu tests"
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r3a_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r3a_tests.log
run() { # tag, args...
  tag=$1; shift
  timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/r3a_$tag.json 2> gpurun_out/r3a_$tag.err; rc=$?
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3a_$tag.json")); r=j["roofline"]; c=j["clocks"]
    print("$tag rc=$rc ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"kernel GB/s",round(r["achieved"]),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),r["kernel"][:12],"clk",c.get("sm_mhz"),c.get("reasons"),"fb",j["check"]["fallback_queries"] if j.get("check") else None, {k:v for k,v in (j.get("check") or {}).items() if v is not True and k!="fallback_queries"})
except Exception as e:
    print("$tag rc=$rc FAILED", e); print(open("gpurun_out/r3a_$tag.err").read()[-1500:])
PY
}
S=4829565
run ts_4p8   --rows $S --steps 40 --variant 2
run qs_4p8   --rows $S --steps 40 --variant 3
run qs6_4p8  --rows $S --steps 40 --variant 3 --opt qs_resident_kb=6
run qs3_4p8  --rows $S --steps 40 --variant 3 --opt qs_resident_kb=3
run qsr_4p8  --rows $S --steps 40 --variant 1
run qs_q2_4p8 --rows $S --steps 40 --variant 3 --opt qs_q_stages=2
run qs_q4_4p8 --rows $S --steps 40 --variant 3 --opt qs_q_stages=4
run ts_4p8_aniso --rows $S --steps 40 --variant 2 --data aniso
run qs_4p8_aniso --rows $S --steps 40 --variant 3 --data aniso
run ts_4p8_aniso_nocenter --rows $S --steps 5 --variant 2 --data aniso --opt center=0
run ts_38   --variant 2
run qs_38   --variant 3
run qs6_38  --variant 3 --opt qs_resident_kb=6
run qs_38_aniso --variant 3 --data aniso
