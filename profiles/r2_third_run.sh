#!/bin/bash
mkdir -p gpurun_out
python profiles/suite_contention.py 50000 Bukin06 Thurber Cola > gpurun_out/r2_contention_smem.log 2>&1
SDA_LS_SMEM_MAXDIM=0 python profiles/suite_contention.py 50000 Bukin06 >> gpurun_out/r2_contention_smem.log 2>&1
(time python -m pytest tests -m gpu -q -x -k "not whole_catalogue") > gpurun_out/r2_pytest_gpu_3.log 2>&1
(time python -m sdaopt_b200.benchmark.workflow --nb-runs 100 --output-folder gpurun_out/DATA_r2b) > gpurun_out/r2_suite_1gpu_b.log 2>&1
rm -f gpurun_out/DATA_r2b/*.data
cat gpurun_out/r2_contention_smem.log; tail -n 4 gpurun_out/r2_pytest_gpu_3.log; tail -n 4 gpurun_out/r2_suite_1gpu_b.log
