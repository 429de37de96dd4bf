#!/bin/bash
# round 2, call F: TS with 8 epilogue warps (tests!), QS resident K-blocks x query-ring depth, LSU-path L2 prefetch, loader threads
mkdir -p gpurun_out
echo "=== gpu tests"
timeout 500 python -m pytest tests -q -m gpu -x --timeout 120 > gpurun_out/r3f_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r3f_tests.log
run() { # tag, args...
  tag=$1; shift
  timeout 200 python bench.py --no-cpu-baseline --no-oracle-check "$@" > gpurun_out/r3f_$tag.json 2> gpurun_out/r3f_$tag.err; rc=$?
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3f_$tag.json")); r=j["roofline"]; c=j["clocks"]; s=j.get("sustained") or {}
    sc=(s.get("clocks") or {})
    print("$tag rc=$rc ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"kern GB/s",round(r["achieved"]),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel",round(r["select_kernels_ms_per_step"],3),r["kernel"][5:8],"clk",c.get("sm_mhz"),"| sus",round(s.get("ms_per_step",0),3),"GB/s",round(s.get("streamed_gbs_per_gpu",0)),"clk",sc.get("sm_mhz"),sc.get("power_w_median"),"| fb",(j.get("check") or {}).get("fallback_queries"))
except Exception as e:
    print("$tag rc=$rc FAILED", e); print(open("gpurun_out/r3f_$tag.err").read()[-1500:])
PY
}
S=4829565
run ts_4p8_nq173  --rows $S --steps 40 --sustain-seconds 0 --variant 2
run ts_4p8_nq208  --rows $S --steps 40 --sustain-seconds 0 --variant 2 --nq 208
run ts_4p8_nq256  --rows $S --steps 40 --sustain-seconds 0 --variant 2 --nq 256
run ts_4p8_k1000  --rows $S --steps 20 --sustain-seconds 0 --variant 2 --k 1000
for R in 12 10 9 8 6; do
for Q in 2 3; do
run qs_R${R}_q$Q  --rows $S --steps 40 --sustain-seconds 0 --opt qs_resident_kb=$R --opt qs_q_stages=$Q
done; done
for pf in 1 2 4; do
run qsr_lsupf$pf  --rows $S --steps 40 --sustain-seconds 0 --opt l2_prefetch=$pf
done
run qs_R8_q2_38 --opt qs_resident_kb=8 --opt qs_q_stages=2
run qsr_38 
run qsr_38_lsupf2 --opt l2_prefetch=2
echo "=== loader"
timeout 300 python tools/load_bench.py 1 2000000 > gpurun_out/r3f_load_1gpu.json 2> gpurun_out/r3f_load_1gpu.err; echo "rc=$?"; cat gpurun_out/r3f_load_1gpu.json; tail -3 gpurun_out/r3f_load_1gpu.err
