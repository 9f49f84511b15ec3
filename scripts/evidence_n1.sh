#!/bin/bash
# single-GPU evidence batch (under gpurun): plain runs, then ONE ncu pass (launch list of the default bench command)
timeout 300 python bench.py --steps 20 --warmup 5 --dim 64 --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r2_n1_d64.json
timeout 200 python scripts/random_row_peak.py > gpurun_out/r2_random_rows.json 2>/dev/null
timeout 300 python scripts/bench_models.py 2>gpurun_out/r2_models.err > gpurun_out/r2_models.json
timeout 200 python scripts/time_small_epoch.py > gpurun_out/r2_small_epoch.txt 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eval-users 32768 --no-eval-full > gpurun_out/r2_plain_for_ncu.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_slarge.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eval-users 32768 --no-eval-full > gpurun_out/r2_ncu_launches.log 2>&1
tail -2 gpurun_out/r2_models.err; cat gpurun_out/r2_small_epoch.txt | tail -2; cat gpurun_out/r2_random_rows.json; wc -l gpurun_out/r2_launches_slarge.csv
