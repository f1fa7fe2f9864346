#!/bin/bash
mkdir -p gpurun_out
rm -rf gpurun_out/DATA_r2g
(time python -m sdaopt_b200.benchmark.workflow --nb-runs 100 --output-folder gpurun_out/DATA_r2g) > gpurun_out/r2_suite_1gpu_g.log 2>&1
rm -f gpurun_out/DATA_r2g/*.data
tail -n 4 gpurun_out/r2_suite_1gpu_g.log
diff <(cut -d, -f1-7 gpurun_out/DATA_r2g/results.csv) <(cut -d, -f1-7 profiles/r2_suite_full_results.csv) | head -5; echo "csv diff done"
(time timeout 600 python -m pytest tests -m gpu -q -x -k "suite") > gpurun_out/r2_pytest_15.log 2>&1
tail -n 3 gpurun_out/r2_pytest_15.log
