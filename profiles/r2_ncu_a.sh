#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-suite --no-shapes > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-suite --no-shapes > gpurun_out/ncu_bench.log 2>&1
tail -n 2 gpurun_out/ncu_bench.log; wc -l gpurun_out/r2_launches_bench.csv
SDA_PHASED=1 python profiles/prof_run.py --chains 65536 --maxiter 200 --reps 1 > gpurun_out/plain_prof.log 2>&1 &&
SDA_PHASED=1 ncu --set full --clock-control none --import-source on -k regex:sda_anneal -s 20 -c 1 -f -o gpurun_out/r2_prof_anneal python profiles/prof_run.py --chains 65536 --maxiter 200 --reps 1 > gpurun_out/ncu_anneal.log 2>&1
tail -n 2 gpurun_out/ncu_anneal.log; du -sh gpurun_out
