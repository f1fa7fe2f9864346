#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_pytest_gpu_full.log 2>&1
python profiles/suite_contention.py 50000 Bukin06 > gpurun_out/r2_contention.log 2>&1
SDA_GROUP=8 python profiles/suite_contention.py 50000 Bukin06 >> gpurun_out/r2_contention.log 2>&1
SDA_GROUP=2 python profiles/suite_contention.py 50000 Bukin06 >> gpurun_out/r2_contention.log 2>&1
python profiles/suite_contention.py 50000 Cola Thurber >> gpurun_out/r2_contention.log 2>&1
tail -n 5 gpurun_out/r2_pytest_gpu_full.log; cat gpurun_out/r2_contention.log
