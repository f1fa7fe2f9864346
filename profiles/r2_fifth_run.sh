#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -q -x -k "phase_clock") > gpurun_out/r2_pytest_sync.log 2>&1
export SDA_CONT_NS=148,1184,2368,4736
for m in 0 1; do
SDA_PERSIST_SYNC=$m timeout 300 python profiles/suite_contention.py 50000 Bukin06 Thurber Cola > gpurun_out/r2_contention_sync$m.log 2>&1
done
(time timeout 600 python -m sdaopt_b200.benchmark.workflow --nb-runs 100 --output-folder gpurun_out/DATA_r2c) > gpurun_out/r2_suite_1gpu_c.log 2>&1
rm -rf gpurun_out/DATA_r2c
tail -n 3 gpurun_out/r2_pytest_sync.log; cat gpurun_out/r2_contention_sync0.log gpurun_out/r2_contention_sync1.log | cut -c1-120; tail -n 4 gpurun_out/r2_suite_1gpu_c.log
