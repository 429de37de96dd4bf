#!/bin/bash
# same-box A/B: refresher pacing with the two-level histogram read
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_gpu_all.log
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
run 4829565 40 adapt2000_$round
run 4829565 40 adapt500_$round --tighten 500
run 4829565 40 fix2000_$round --opt tighten_adaptive=0 --tighten 2000
run 4829565 40 fix10000_$round --opt tighten_adaptive=0 --tighten 10000
run 4829565 40 fix30000_$round --opt tighten_adaptive=0 --tighten 30000
done
run 38636520 20 adapt_38
run 38636520 20 fix10000_38 --opt tighten_adaptive=0 --tighten 10000
