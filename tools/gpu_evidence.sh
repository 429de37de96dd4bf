#!/bin/bash
# Evidence for profiles/ on ONE GPU (round 2): full GPU test-suite, smoke, headline bench with default flags (CPU
# baseline, full-size oracle check, sustained record), reference arm, per-GPU shard of the 8-GPU case, BASELINE
# config 2, ncu launch lists, ncu --set full of the dominant kernel at both sizes, engine sweeps.
# Outputs under gpurun_out/ (copied to profiles/ by tools/collect_profiles.py r02).
mkdir -p gpurun_out
echo "=== full gpu test-suite"
timeout 600 python -m pytest tests -q -m gpu --timeout 300 -rs > gpurun_out/t_gpu_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gpu_all.log
echo "=== smoke()"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_reference.json
echo "=== headline bench (default flags)"
timeout 600 python bench.py > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err; echo "rc=$?"; tail -3 gpurun_out/bench_headline.err; cut -c1-3000 gpurun_out/bench_headline.json
echo "=== per-GPU shard of the 8-GPU case on one GPU"
timeout 300 python bench.py --rows 4829565 --steps 40 --no-cpu-baseline > gpurun_out/bench_4p8M.json 2> gpurun_out/bench_4p8M.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_4p8M.json
echo "=== BASELINE config 2 (8.8M rows, 173 queries, top-1000, single GPU)"
timeout 300 python bench.py --config c2 --steps 20 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_c2.json
echo "=== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_headline.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-check --sustain-seconds 0 > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_4p8M.csv python bench.py --rows 4829565 --steps 2 --warmup 3 --no-cpu-baseline --no-check --sustain-seconds 0 > gpurun_out/ncu_list2.log 2>&1; echo "rc=$?"
echo "=== ncu --set full, dominant kernel (the one scoring launch of a step)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_qs -s 3 -c 1 -o gpurun_out/prof_qs_headline python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check --sustain-seconds 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_qs -s 3 -c 1 -o gpurun_out/prof_qs_4p8M python bench.py --rows 4829565 --steps 1 --warmup 3 --no-cpu-baseline --no-check --sustain-seconds 0 > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full2.log
echo "=== engine sweeps (AUTO vs forced engines)"
timeout 400 python tools/sweep.py 8841823 100 gpurun_out/sweep_8p8M_k100.json 2>&1 | tail -12 | cut -c1-330
timeout 400 python tools/sweep.py 8841823 1000 gpurun_out/sweep_8p8M_k1000.json 1,2,4,8,173,256 2>&1 | tail -6 | cut -c1-330
