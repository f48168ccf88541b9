#!/bin/bash
# One-GPU evidence pass (run under gpurun from the repo root): full GPU suite, headline bench (+ CPU baseline), reference arm,
# ncu launch list of a short bench run, ncu --set full of the two headline kernels (exported as CSV: the .ncu-rep stays on the box).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -8 > gpurun_out/r2_gputest.txt
timeout 600 python bench.py 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/r2_bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench_err.txt | tail -1 > gpurun_out/r2_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_ncu.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trigemm_kernel|kstar_digits_kernel|tail_kernel" -c 6 \
  -o /tmp/r2_oz -f python tools/prof_oz.py > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/r2_oz.ncu-rep --page raw --csv > gpurun_out/r2_oz_raw.csv 2>>gpurun_out/ncu_full.log
python tools/ncu_summary.py gpurun_out/r2_oz_raw.csv > gpurun_out/r2_int8_engine_ncu_summary_tables.md 2>>gpurun_out/ncu_full.log
tail -3 gpurun_out/r2_gputest.txt; cat gpurun_out/r2_bench_final.json gpurun_out/r2_bench_reference.json
