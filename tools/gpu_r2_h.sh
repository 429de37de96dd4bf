#!/bin/bash
# bootstrap_select_kernel: parity + launch list + bench
mkdir -p gpurun_out
nvidia-smi -q -d POWER | grep -i -E "power limit|power draw" | head -6
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu_all.log
for cfg in "4829565 40" "38636520 20"; do
set -- $cfg
timeout 900 python bench.py --rows $1 --steps $2 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "rc=$?"; tail -2 gpurun_out/bench_$1.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_$1.json")); r=j["roofline"]; c=j["clocks"]
print("rows $1: ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"kernel GB/s",round(r["achieved"]),"ms/launch",round(r["ms_per_launch"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"launches/step",j["gpu_launches"]/j["steps"],"clk",c,"check",j["check"])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_4p8M.csv python bench.py --rows 4829565 --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_list.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_4p8M.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-6:]:
    print(r[4][:60], r[-1])
PY
