#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/ab_bisect2.log; rm -f $out
for r in 1 2; do for v in r1 mb3 xa xbb xccc; do
  echo "== $v" >> $out
  SDA_LIB_PATH=$PWD/build/ab/lib_$v.so timeout 300 python profiles/prof_run.py --chains 65536 --maxiter 400 --pure-sa --reps 2 2>&1 | tail -n 1 >> $out
done; done
for v in r1 xa xccc; do
  echo "== $v LS on maxiter 1000" >> $out
  SDA_LIB_PATH=$PWD/build/ab/lib_$v.so timeout 300 python profiles/prof_run.py --chains 65536 --maxiter 1000 --reps 2 2>&1 | tail -n 1 >> $out
done
cat $out
