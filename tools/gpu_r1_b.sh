#!/bin/bash
# Round 1, session B: headline config on one GPU, per-GPU-size runs, ncu evidence.
mkdir -p gpurun_out
echo "=== misc tests again"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden" > gpurun_out/t_misc2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_misc2.log
echo "=== bench headline (38.6M rows, 1 GPU)"
timeout 900 python bench.py --steps 10 > gpurun_out/bench_full_1gpu.json 2> gpurun_out/bench_full_1gpu.err; echo "rc=$?"; cat gpurun_out/bench_full_1gpu.json; tail -5 gpurun_out/bench_full_1gpu.err
for g in 16 4; do
echo "=== bench 4.83M rows growth=$g"
timeout 600 python bench.py --rows 4829565 --steps 20 --no-cpu-baseline --growth $g > gpurun_out/bench_4p8M_g$g.json 2> gpurun_out/bench_4p8M_g$g.err; echo "rc=$?"; cat gpurun_out/bench_4p8M_g$g.json; tail -3 gpurun_out/bench_4p8M_g$g.err
done
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_4p8M.csv python bench.py --rows 4829565 --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
grep -c umma gpurun_out/launches_4p8M.csv
echo "=== ncu full on the scoring kernel"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:umma_score -s 6 -c 3 -o gpurun_out/prof_umma_4p8M python bench.py --rows 4829565 --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
