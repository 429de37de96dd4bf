#!/bin/bash
mkdir -p gpurun_out
for cfg in "173 40" "64 40" "16 40"; do
set -- $cfg
timeout 900 python bench.py --rows 38636520 --nq $1 --steps $2 --no-cpu-baseline --no-check > gpurun_out/bench_pw_$1.json 2> gpurun_out/bench_pw_$1.err; echo "rc=$?"; tail -2 gpurun_out/bench_pw_$1.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_pw_$1.json")); r=j["roofline"]
print("nq",$1,"ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"kernel GB/s",round(r["achieved"]),"clocks",j["clocks"])
PY
done
nvidia-smi --query-gpu=power.limit,power.max_limit,clocks.max.sm,clocks.max.mem --format=csv
