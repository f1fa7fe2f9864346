#!/bin/bash
# same-box A/B of the builds under build/ab/: headline shape, pure SA (maxiter 100) and with local search (maxiter 300)
mkdir -p gpurun_out; out=gpurun_out/${AB_OUT:-ab_variants.log}; rm -f $out
for lib in ${AB_LIBS:-$(ls build/ab/lib_*.so)}; do
  echo "== $lib" >> $out
  SDA_LIB_PATH=$PWD/$lib timeout 120 python profiles/prof_run.py --chains 65536 --maxiter 100 --pure-sa --reps 3 2>&1 | tail -n 1 >> $out
  SDA_LIB_PATH=$PWD/$lib timeout 200 python profiles/prof_run.py --chains 65536 --maxiter 300 --reps 2 2>&1 | tail -n 1 >> $out
done
cat $out
