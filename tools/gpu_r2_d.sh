#!/bin/bash
# async API + packed exchange: parity, N=1 benches (short and long steps)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_gpu_all.log
for cfg in "4829565 40" "9659130 30" "38636520 20"; do
set -- $cfg
timeout 900 python bench.py --rows $1 --steps $2 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "rc=$?"; tail -2 gpurun_out/bench_$1.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_$1.json")); r=j["roofline"]; c=j["clocks"]
print("rows $1: ms/step",round(j["ms_per_step"],3),"q/s",round(j["value"]),"e2e",round(j["e2e"]["value"]),"kernel GB/s",round(r["achieved"]),"ms/launch",round(r["ms_per_launch"],3),"sel ms",round(r["select_kernels_ms_per_step"],3),"launches/step",j["gpu_launches"]/j["steps"],"clk",c.get("sm_mhz"),c.get("reasons"),"check",j["check"])
PY
done
