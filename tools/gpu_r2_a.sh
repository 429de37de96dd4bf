#!/bin/bash
# histogram tightening: parity + same-box A/B against the phase schedule
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_gpu_all.log
for cfg in "4829565 40 400" "4829565 40 0" "4829565 40 100" "4829565 40 2000" "9659130 30 400" "9659130 30 0" "38636520 20 400" "38636520 20 0"; do
set -- $cfg
timeout 900 python bench.py --rows $1 --steps $2 --tighten $3 --no-cpu-baseline > gpurun_out/bench_$1_t$3.json 2> gpurun_out/bench_$1_t$3.err; echo "rc=$?"; tail -2 gpurun_out/bench_$1_t$3.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_$1_t$3.json")); r=j["roofline"]; c=j["clocks"]
print("rows $1 tighten $3: ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"kernel GB/s",round(r["achieved"]),"ms/launch",round(r["ms_per_launch"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"launches/step",j["gpu_launches"]/j["steps"],"clk",c.get("sm_mhz"),c.get("reasons"),"fallback",j["check"]["fallback_queries"],"ok",j["check"]["tensor_engine_equals_simt_engine_4q"], j["check"]["max_rel_err_vs_host_fp64_rescoring"])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_4p8M.csv python bench.py --rows 4829565 --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_list.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_4p8M.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-9:]:
    print(r[4][:60], r[-1])
PY
