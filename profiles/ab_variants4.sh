#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/ab_variants5.log; rm -f $out
run() { echo "== $1 $2" >> $out
  env $2 SDA_LIB_PATH=$PWD/$1 timeout 120 python profiles/prof_run.py --chains 65536 --maxiter 100 --pure-sa --reps 3 2>&1 | tail -n 1 >> $out
  env $2 SDA_LIB_PATH=$PWD/$1 timeout 200 python profiles/prof_run.py --chains 65536 --maxiter 300 --reps 2 2>&1 | tail -n 1 >> $out; }
run build/ab/lib_prev_mb4.so X=1
run build/ab/lib_cur4.so X=1
run sdaopt_b200/libsdaopt_b200.so X=1
run build/ab/lib_prev_mb4.so X=2
run build/ab/lib_cur4.so X=2
cat $out
timeout 300 python profiles/shapes_bench.py --quick --reps 1 2>&1 | grep -v "^{"
