#!/bin/bash
# no-bootstrap single-launch mode: parity + same-box A/B against the dense bootstrap
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu_all.log
run() { # rows steps tag opts...
rows=$1; steps=$2; tag=$3; shift 3
timeout 900 python bench.py --rows $rows --steps $steps --no-cpu-baseline "$@" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; tail -2 gpurun_out/ab_$tag.err
python - <<PY
import json
j=json.load(open("gpurun_out/ab_$tag.json")); r=j["roofline"]; c=j["clocks"]
print("$tag rows $rows: ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"score ms/step",round(r["score_kernel_share_of_step"]*j["ms_per_step"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"launches",j["gpu_launches"]/j["steps"],"clk",c.get("sm_mhz"),c.get("sm_mhz_min"),c.get("reasons"),"chk",j["check"]["tensor_engine_equals_simt_engine_4q"],j["check"]["fallback_queries"])
PY
}
for round in 1 2; do
run 4829565 40 nb_4p8_$round
run 4829565 40 bs_4p8_$round --opt bootstrap=1
done
run 38636520 20 nb_38
run 38636520 20 bs_38 --opt bootstrap=1
run 8841823 30 nb_k1000 --k 1000
run 8841823 30 bs_k1000 --k 1000 --opt bootstrap=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_4p8M.csv python bench.py --rows 4829565 --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_list.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_4p8M.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-4:]:
    print(r[4][:60], r[-1])
PY
