#!/bin/bash
# round-2 evidence run on one B200: parity pins, GPU tests, smoke, bench (both arms), suite, config 3/4 table
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_final_gpu.txt
python profiles/record_parity_pins.py > gpurun_out/r2_final_pins.log 2>&1
cp gpurun_out/gpu_pins.json tests/golden/gpu_pins.json
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2_final_pytest_gpu.log 2>&1
tail -n 4 gpurun_out/r2_final_pytest_gpu.log
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2_final_smoke.log 2>&1
tail -n 4 gpurun_out/r2_final_smoke.log
(time python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
(time python bench.py) > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
tail -n 4 gpurun_out/r2_final_bench.err
rm -rf gpurun_out/DATA_r2f
(time python -m sdaopt_b200.benchmark.workflow --nb-runs 100 --output-folder gpurun_out/DATA_r2f) > gpurun_out/r2_final_suite_1gpu.log 2>&1
rm -f gpurun_out/DATA_r2f/*.data
tail -n 4 gpurun_out/r2_final_suite_1gpu.log
timeout 900 python profiles/shapes_bench.py --reps 1 > gpurun_out/r2_final_shapes.jsonl 2> gpurun_out/r2_final_shapes.txt
cat gpurun_out/r2_final_shapes.txt
