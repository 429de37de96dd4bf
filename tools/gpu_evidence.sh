#!/bin/bash
# Evidence for profiles/: full GPU test-suite, headline bench, ncu launch list (same command),
# ncu --set full of the dominant kernel.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
echo "=== full gpu test-suite"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gpu_all.log
echo "=== smoke()"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log
echo "=== headline bench (default flags)"
timeout 900 python bench.py > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err; echo "rc=$?"; tail -3 gpurun_out/bench_headline.err; cat gpurun_out/bench_headline.json
echo "=== reference arm"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cat gpurun_out/bench_reference.json
echo "=== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_headline.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "=== ncu --set full, dominant kernel (last, largest phase)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:umma_score -s 11 -c 1 -o gpurun_out/prof_umma_headline python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
