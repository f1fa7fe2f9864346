#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/ab_variants6.log; rm -f $out
run() { echo "== $1 $2" >> $out
  env $2 SDA_LIB_PATH=$PWD/$1 timeout 200 python profiles/prof_run.py --chains 65536 --maxiter 300 --reps 2 2>&1 | tail -n 1 >> $out; }
run build/ab/lib_cur.so X=1
run build/ab/lib_lsb3.so X=1
run build/ab/lib_lsb4.so X=1
run build/ab/lib_cur.so SDA_STEPS_PER_ROUND=4
run build/ab/lib_cur.so SDA_STEPS_PER_ROUND=6
run build/ab/lib_lsb3.so SDA_STEPS_PER_ROUND=4
run build/ab/lib_cur.so SDA_LS_SYNC=0
cat $out
