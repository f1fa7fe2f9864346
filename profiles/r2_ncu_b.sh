#!/bin/bash
mkdir -p gpurun_out
SDA_PHASED=1 python profiles/prof_run.py --chains 65536 --maxiter 200 --reps 1 > gpurun_out/plain_prof.log 2>&1 &&
SDA_PHASED=1 ncu --set full --clock-control none --import-source on -k regex:sda_ls_phase -s 5 -c 1 -f -o gpurun_out/r2_prof_ls python profiles/prof_run.py --chains 65536 --maxiter 200 --reps 1 > gpurun_out/ncu_ls.log 2>&1
tail -n 2 gpurun_out/ncu_ls.log; du -sh gpurun_out
