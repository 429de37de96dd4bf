#!/bin/bash
# same-box A/B of the tightening pause (two rounds to expose drift)
mkdir -p gpurun_out
for round in 1 2; do
for t in 0 2000 20000 100000; do
timeout 900 python bench.py --rows 4829565 --steps 60 --tighten $t --no-cpu-baseline --no-check > gpurun_out/ab_$t.json 2> gpurun_out/ab_$t.err; tail -1 gpurun_out/ab_$t.err
python - <<PY
import json
j=json.load(open("gpurun_out/ab_$t.json")); r=j["roofline"]; c=j["clocks"]
print("round $round tighten $t: ms/step",round(j["ms_per_step"],3),"kernel ms/step",round(r["ms_per_launch"]*j["gpu_launches"]/j["steps"]/8*0+r["score_kernel_share_of_step"]*j["ms_per_step"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"launches/step",j["gpu_launches"]/j["steps"],"clk",c.get("sm_mhz"),c.get("sm_mhz_min"),c.get("reasons"))
PY
done
done
for t in 0 2000 20000; do
timeout 900 python bench.py --rows 38636520 --steps 30 --tighten $t --no-cpu-baseline --no-check > gpurun_out/ab38_$t.json 2> gpurun_out/ab38_$t.err; tail -1 gpurun_out/ab38_$t.err
python - <<PY
import json
j=json.load(open("gpurun_out/ab38_$t.json")); r=j["roofline"]; c=j["clocks"]
print("38.6M tighten $t: ms/step",round(j["ms_per_step"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"clk",c.get("sm_mhz"),c.get("sm_mhz_min"),c.get("reasons"))
PY
done
