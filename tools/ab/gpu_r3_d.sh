#!/bin/bash
# round 2, call D: first-tile protocol + refresher priority: tests, fallback counts, batch-size/variant sweep
mkdir -p gpurun_out
echo "=== gpu tests"
timeout 400 python -m pytest tests -q -m gpu -x --timeout 120 > gpurun_out/r3d_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r3d_tests.log
run() { # tag, args...
  tag=$1; shift
  timeout 150 python bench.py --no-cpu-baseline --no-oracle-check "$@" > gpurun_out/r3d_$tag.json 2> gpurun_out/r3d_$tag.err; rc=$?
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3d_$tag.json")); r=j["roofline"]; c=j["clocks"]; s=j.get("sustained") or {}
    sc=(s.get("clocks") or {})
    print("$tag rc=$rc ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"kern GB/s",round(r["achieved"]),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel",round(r["select_kernels_ms_per_step"],3),r["kernel"][5:8],"clk",c.get("sm_mhz"),"| sus",round(s.get("ms_per_step",0),3),"GB/s",round(s.get("streamed_gbs_per_gpu",0)),"clk",sc.get("sm_mhz"),sc.get("power_w_median"),"| fb",(j.get("check") or {}).get("fallback_queries"))
except Exception as e:
    print("$tag rc=$rc FAILED", e); print(open("gpurun_out/r3d_$tag.err").read()[-1500:])
PY
}
S=4829565
for nq in 16 64 128 173 208; do
run ts_4p8_nq$nq   --rows $S --steps 40 --variant 2 --nq $nq --sustain-seconds 0
run qsr_4p8_nq$nq  --rows $S --steps 40 --variant 3 --nq $nq --sustain-seconds 0
run qs0_4p8_nq$nq  --rows $S --steps 40 --variant 3 --nq $nq --sustain-seconds 0 --opt qs_resident_kb=0
done
run qsr_4p8_k1000  --rows $S --steps 20 --variant 3 --k 1000 --sustain-seconds 0
run ts_4p8_k1000   --rows $S --steps 20 --variant 2 --k 1000 --sustain-seconds 0
run qsr_4p8_aniso  --rows $S --steps 40 --variant 3 --data aniso --sustain-seconds 0
for nq in 16 64 173; do
run qsr_38_nq$nq --variant 3 --nq $nq
done
run ts_38_nq64 --variant 2 --nq 64
