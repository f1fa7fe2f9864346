#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r2_anneal_memo_ab.log; rm -f $out
(timeout 600 python -m pytest tests -m gpu -q -x -k "memo or philox or trajectory or phased") > gpurun_out/r2_pytest_amemo.log 2>&1
tail -n 3 gpurun_out/r2_pytest_amemo.log
for m in 0 1 0 1; do
echo "== SDA_ANNEAL_MEMO=$m" >> $out
SDA_ANNEAL_MEMO=$m timeout 120 python profiles/prof_run.py --chains 65536 --maxiter 100 --pure-sa --reps 3 2>&1 | tail -n 1 >> $out
SDA_ANNEAL_MEMO=$m timeout 200 python profiles/prof_run.py --chains 65536 --maxiter 300 --reps 2 2>&1 | tail -n 1 >> $out
done
cat $out
