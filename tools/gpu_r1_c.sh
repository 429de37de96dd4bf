#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest umma + misc"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not scan_f32 and not scan_exact" --maxfail=6 > gpurun_out/t_umma.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_umma.log
for rows in 4829565 38636520; do
echo "=== bench rows=$rows"
timeout 900 python bench.py --rows $rows --steps 10 --no-cpu-baseline > gpurun_out/bench_$rows.json 2> gpurun_out/bench_$rows.err; echo "rc=$?"; tail -3 gpurun_out/bench_$rows.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_$rows.json"))
r=j["roofline"]
print("value",j["value"],"ms/step",j["ms_per_step"],"achieved",r["achieved"],"frac",r["frac"],"ms/launch",r["ms_per_launch"],"share",r["score_kernel_share_of_step"],"select_ms",r["select_kernels_ms_per_step"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],j["check"])
PY
done
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_4p8M.csv python bench.py --rows 4829565 --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_4p8M.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-22:]:
    print(r[4][:60], r[-1], r[-2])
PY
