#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r2_run14.log; rm -f $out
(time timeout 1500 python -m pytest tests -m gpu -q -x -k "not whole_catalogue") > gpurun_out/r2_pytest_14.log 2>&1
tail -n 4 gpurun_out/r2_pytest_14.log
timeout 120 python profiles/prof_run.py --chains 65536 --maxiter 400 --pure-sa --reps 2 2>&1 | tail -n 1 >> $out
timeout 300 python profiles/prof_run.py --chains 65536 --maxiter 1000 --reps 2 2>&1 | tail -n 1 >> $out
SDA_LIB_PATH=$PWD/build/ab/lib_r1.so timeout 300 python profiles/prof_run.py --chains 65536 --maxiter 1000 --reps 2 2>&1 | tail -n 1 >> $out
timeout 600 python profiles/shapes_bench.py --quick --reps 1 2>&1 | grep -v "^{" >> $out
cat $out
