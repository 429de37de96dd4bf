#!/bin/bash
# round 2, call C: new epilogues (predicate chains; QS with 8 epilogue warps): tests, then TS vs QS again,
# burst (4.83M-row shard) and sustained (38.6M rows), with the full-size oracle check.
mkdir -p gpurun_out
echo "=== gpu tests"
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r3c_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r3c_tests.log
run() { # tag, args...
  tag=$1; shift
  timeout 900 python bench.py --no-cpu-baseline "$@" > gpurun_out/r3c_$tag.json 2> gpurun_out/r3c_$tag.err; rc=$?
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3c_$tag.json")); r=j["roofline"]; c=j["clocks"]; s=j.get("sustained") or {}
    sc=(s.get("clocks") or {})
    print("$tag rc=$rc ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"kern GB/s",round(r["achieved"]),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel",round(r["select_kernels_ms_per_step"],3),r["kernel"][5:8],"clk",c.get("sm_mhz"),c.get("reasons"),"| sustained",round(s.get("ms_per_step",0),3),"GB/s",round(s.get("streamed_gbs_per_gpu",0)),"clk",sc.get("sm_mhz"),sc.get("power_w_median"),"| check",{k:v for k,v in (j.get("check") or {}).items() if v is not True and not k.startswith("oracle_") or k=="oracle_violations"})
except Exception as e:
    print("$tag rc=$rc FAILED", e); print(open("gpurun_out/r3c_$tag.err").read()[-1500:])
PY
}
S=4829565
run ts_4p8   --rows $S --steps 40 --variant 2 --no-oracle-check
run qs_4p8   --rows $S --steps 40 --variant 3
run qs6_4p8  --rows $S --steps 40 --variant 3 --opt qs_resident_kb=6 --no-oracle-check
run qsr_4p8  --rows $S --steps 40 --variant 1 --no-oracle-check
run qs_q4_4p8 --rows $S --steps 40 --variant 3 --opt qs_q_stages=4 --no-oracle-check
run qs_4p8_aniso --rows $S --steps 40 --variant 3 --data aniso
run qsr_4p8_aniso --rows $S --steps 40 --variant 1 --data aniso --no-oracle-check
run ts_4p8_aniso --rows $S --steps 40 --variant 2 --data aniso --no-oracle-check
run ts_38   --variant 2 --no-oracle-check
run qs_38   --variant 3
run qs6_38  --variant 3 --opt qs_resident_kb=6 --no-oracle-check
run qsr_38  --variant 1 --no-oracle-check
run qs_38_nq16  --variant 3 --nq 16 --no-oracle-check
run qs_38_nq64  --variant 3 --nq 64 --no-oracle-check
