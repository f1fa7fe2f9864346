#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/r2_memo_ab.log; rm -f $out
(timeout 600 python -m pytest tests -m gpu -q -x -k "memoised or local_search or phased or block_synch") > gpurun_out/r2_pytest_memo.log 2>&1
tail -n 3 gpurun_out/r2_pytest_memo.log
for m in 0 1; do
echo "== SDA_LS_MEMO=$m" >> $out
SDA_LS_MEMO=$m timeout 200 python profiles/prof_run.py --chains 65536 --maxiter 300 --reps 2 2>&1 | tail -n 1 >> $out
SDA_LS_MEMO=$m timeout 400 python profiles/shapes_bench.py --quick --only rastrigin --reps 1 2>&1 | grep -v "^{" >> $out
done
cat $out
