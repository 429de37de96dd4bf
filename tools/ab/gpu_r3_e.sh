#!/bin/bash
# round 2, call E: L2 prefetch distance for the QS kernel (burst + sustained), ncu of the current QS kernel,
# loader throughput, the new tests.
mkdir -p gpurun_out
echo "=== new gpu tests"
timeout 300 python -m pytest tests -q -m gpu -x --timeout 120 -k "rank_dedup or flat_file or resident or config1 or common_mean" > gpurun_out/r3e_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r3e_tests.log
run() { # tag, args...
  tag=$1; shift
  timeout 200 python bench.py --no-cpu-baseline --no-oracle-check "$@" > gpurun_out/r3e_$tag.json 2> gpurun_out/r3e_$tag.err; rc=$?
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r3e_$tag.json")); r=j["roofline"]; c=j["clocks"]; s=j.get("sustained") or {}
    sc=(s.get("clocks") or {})
    print("$tag rc=$rc ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"kern GB/s",round(r["achieved"]),"frac",round(r["frac"],3),"ms/launch",round(r["ms_per_launch"],3),"sel",round(r["select_kernels_ms_per_step"],3),r["kernel"][5:8],"clk",c.get("sm_mhz"),"| sus",round(s.get("ms_per_step",0),3),"GB/s",round(s.get("streamed_gbs_per_gpu",0)),"clk",sc.get("sm_mhz"),sc.get("power_w_median"),"| fb",(j.get("check") or {}).get("fallback_queries"))
except Exception as e:
    print("$tag rc=$rc FAILED", e); print(open("gpurun_out/r3e_$tag.err").read()[-1500:])
PY
}
S=4829565
for pf in 0 1 2 3; do
run qsr_4p8_pf$pf  --rows $S --steps 40 --sustain-seconds 0 --opt l2_prefetch=$pf
done
run qs0_4p8_pf0  --rows $S --steps 40 --sustain-seconds 0 --opt l2_prefetch=0 --opt qs_resident_kb=0
run qs0_4p8_pf2  --rows $S --steps 40 --sustain-seconds 0 --opt l2_prefetch=2 --opt qs_resident_kb=0
run qsr_4p8_nq128_pf0  --rows $S --steps 40 --nq 128 --sustain-seconds 0 --opt l2_prefetch=0
run qsr_4p8_nq128_pf2  --rows $S --steps 40 --nq 128 --sustain-seconds 0 --opt l2_prefetch=2
run qsr_4p8_nq192  --rows $S --steps 40 --nq 192 --sustain-seconds 0
run qsr_38_pf0 --opt l2_prefetch=0
run qsr_38_pf1 --opt l2_prefetch=1
run qsr_38_pf2 --opt l2_prefetch=2
echo "=== ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_qs -s 3 -c 1 -o gpurun_out/r3e_prof_qsr_4p8 python bench.py --rows $S --steps 1 --warmup 3 --no-cpu-baseline --no-check --sustain-seconds 0 > gpurun_out/r3e_ncu1.log 2>&1; echo "rc=$?"
echo "=== loader"
timeout 300 python tools/load_bench.py 1 2000000 > gpurun_out/r3e_load_1gpu.json 2> gpurun_out/r3e_load_1gpu.err; echo "rc=$?"; cat gpurun_out/r3e_load_1gpu.json; tail -3 gpurun_out/r3e_load_1gpu.err
df -h /dev/shm /tmp | tail -2; nproc; free -g | head -2
