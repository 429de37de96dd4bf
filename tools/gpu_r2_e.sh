#!/bin/bash
# same-box A/B: data-dependent margin vs worst-case margin (two rounds)
mkdir -p gpurun_out
run() { # rows steps tag opts...
rows=$1; steps=$2; tag=$3; shift 3
timeout 900 python bench.py --rows $rows --steps $steps --no-cpu-baseline --no-check "$@" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; tail -2 gpurun_out/ab_$tag.err
python - <<PY
import json
j=json.load(open("gpurun_out/ab_$tag.json")); r=j["roofline"]; c=j["clocks"]
print("$tag rows $rows: ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"score ms/step",round(r["score_kernel_share_of_step"]*j["ms_per_step"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"clk",c.get("sm_mhz"),c.get("sm_mhz_min"),c.get("reasons"))
PY
}
for round in 1 2; do
run 4829565 40 new_4p8_$round
run 4829565 40 old_4p8_$round --opt worst_case_margin=1
run 38636520 20 new_38_$round
run 38636520 20 old_38_$round --opt worst_case_margin=1
done
run 8841823 30 k1000_new --k 1000
run 8841823 30 k1000_old --k 1000 --opt worst_case_margin=1
