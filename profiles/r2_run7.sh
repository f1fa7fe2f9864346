#!/bin/bash
mkdir -p gpurun_out
python profiles/record_parity_pins.py > gpurun_out/r2_pins3.log 2>&1
export SDA_CONT_NS=100,1600,4736,18944
timeout 600 python profiles/suite_contention.py 50000 Bukin06 Thurber > gpurun_out/r2_contention_thread.log 2>&1
(time timeout 900 python -m pytest tests -m gpu -q -x -k "not whole_catalogue") > gpurun_out/r2_pytest_gpu_7.log 2>&1
(time timeout 600 python -m sdaopt_b200.benchmark.workflow --nb-runs 100 --output-folder gpurun_out/DATA_r2d) > gpurun_out/r2_suite_1gpu_d.log 2>&1
rm -f gpurun_out/DATA_r2d/*.data
grep -v "^    {" gpurun_out/r2_pins3.log | cut -c1-300 | tail -n 12; cat gpurun_out/r2_contention_thread.log | cut -c1-130; tail -n 6 gpurun_out/r2_pytest_gpu_7.log; tail -n 4 gpurun_out/r2_suite_1gpu_d.log
