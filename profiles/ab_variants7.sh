#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/ab_variants7.log; rm -f $out
for s in 8 4 3 2; do
  echo "== SDA_STEPS_PER_ROUND=$s" >> $out
  SDA_STEPS_PER_ROUND=$s timeout 300 python profiles/prof_run.py --chains 65536 --maxiter 1000 --reps 2 2>&1 | tail -n 1 >> $out
done
cat $out
