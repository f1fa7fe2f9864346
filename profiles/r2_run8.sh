#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x -k "not whole_catalogue") > gpurun_out/r2_pytest_gpu_8.log 2>&1
tail -n 4 gpurun_out/r2_pytest_gpu_8.log
SDA_ROWS_GLOBAL=0 timeout 600 python profiles/shapes_bench.py --quick --reps 1 > gpurun_out/shapes_quick_smemrows.jsonl 2> gpurun_out/shapes_quick_smemrows.txt
cat gpurun_out/shapes_quick_smemrows.txt
timeout 900 python profiles/shapes_bench.py --reps 1 > gpurun_out/shapes_r2.jsonl 2> gpurun_out/shapes_r2.txt
cat gpurun_out/shapes_r2.txt
