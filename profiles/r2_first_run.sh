#!/bin/bash
# round 2, first GPU call: record the parity pins, run the GPU tests, smoke, the suite, one bench step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_first_gpu.txt
(time python profiles/record_parity_pins.py) > gpurun_out/r2_pins.log 2>&1
cp gpurun_out/gpu_pins.json tests/golden/gpu_pins.json
(time python -m pytest tests -m gpu -q -x --deselect "tests/test_gpu_parity.py::test_suite_statistics_match_reference_whole_catalogue") > gpurun_out/r2_pytest_gpu.log 2>&1
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2_smoke.log 2>&1
rm -rf gpurun_out/DATA_r2
(time python -m sdaopt_b200.benchmark.workflow --nb-runs 100 --output-folder gpurun_out/DATA_r2) > gpurun_out/r2_suite_1gpu.log 2>&1
rm -f gpurun_out/DATA_r2/*.data
(time python bench.py --steps 1 --warmup 1) > gpurun_out/r2_bench_first.json 2> gpurun_out/r2_bench_first.err
(time python profiles/suite_vs_reference.py suite_stats_ref_1e6.npz) > gpurun_out/r2_suite_vs_ref_1e6.txt 2>&1
tail -3 gpurun_out/r2_pins.log gpurun_out/r2_pytest_gpu.log gpurun_out/r2_smoke.log gpurun_out/r2_suite_1gpu.log
