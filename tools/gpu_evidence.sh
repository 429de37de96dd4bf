#!/bin/bash
# Evidence for profiles/: full GPU test-suite, smoke, headline bench (default flags), reference arm,
# ncu launch list of the same command, ncu --set full of the dominant kernel, batch-size sweeps.
# Outputs under gpurun_out/ (copied to profiles/ by tools/collect_profiles.py).
mkdir -p gpurun_out
echo "=== full gpu test-suite"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gpu_all.log
echo "=== smoke()"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log
echo "=== reference arm"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cat gpurun_out/bench_reference.json
echo "=== headline bench (default flags)"
timeout 900 python bench.py > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err; echo "rc=$?"; tail -3 gpurun_out/bench_headline.err; cat gpurun_out/bench_headline.json
echo "=== per-GPU shard of the 8-GPU case on one GPU"
timeout 900 python bench.py --rows 4829565 --steps 40 --no-cpu-baseline > gpurun_out/bench_4p8M.json 2> gpurun_out/bench_4p8M.err; echo "rc=$?"; cat gpurun_out/bench_4p8M.json
echo "=== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_headline.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_4p8M.csv python bench.py --rows 4829565 --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_list2.log 2>&1; echo "rc=$?"
echo "=== ncu --set full, dominant kernel (the one scoring launch of a step)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:umma_score -s 3 -c 1 -o gpurun_out/prof_umma_headline python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:umma_score -s 3 -c 1 -o gpurun_out/prof_umma_4p8M python bench.py --rows 4829565 --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full2.log
echo "=== batch-size sweeps (BASELINE configs 2 and 5 shapes)"
timeout 1200 python tools/sweep.py 8841823 100 gpurun_out/sweep_8p8M_k100.json 2>&1 | tail -3
timeout 900 python tools/sweep.py 8841823 1000 gpurun_out/sweep_8p8M_k1000.json 8,173,1024 2>&1 | tail -3
timeout 900 python tools/sweep.py 11100000 100 gpurun_out/sweep_11p1M_k100.json 173,1024,5571 2>&1 | tail -3
