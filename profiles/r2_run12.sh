#!/bin/bash
mkdir -p gpurun_out
export SDA_CONT_NS=13
timeout 300 python profiles/suite_contention.py 100000 Thurber DeVilliersGlasser02 PowerSum Meyer 2>&1 | cut -c1-125 > gpurun_out/r2_long_jobs_alone2.log
cat gpurun_out/r2_long_jobs_alone2.log
python profiles/record_parity_pins.py > gpurun_out/r2_pins4.log 2>&1
cp gpurun_out/gpu_pins.json tests/golden/gpu_pins.json
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r2_pytest_gpu_12.log 2>&1
tail -n 8 gpurun_out/r2_pytest_gpu_12.log
(time python profiles/suite_vs_reference.py suite_stats_ref_1e6.npz) > gpurun_out/r2_suite_vs_ref_1e6.txt 2>&1
tail -n 45 gpurun_out/r2_suite_vs_ref_1e6.txt | cut -c1-150
