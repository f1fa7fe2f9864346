#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/ab_ls_clock.log; rm -f $out
for v in cur w16 w4 ph3; do
  lib=$PWD/build/ab/lib_$v.so; [ $v = cur ] && lib=$PWD/sdaopt_b200/libsdaopt_b200.so
  echo "== $v" >> $out
  SDA_LIB_PATH=$lib timeout 300 python bench.py --steps 1 --warmup 1 --no-suite --no-shapes 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('step %.0f ms' % d['ms_per_step'], [(k['kernel'][:22], round(k['ms_per_step'])) for k in d['roofline']['kernels']])" >> $out
done
cat $out
