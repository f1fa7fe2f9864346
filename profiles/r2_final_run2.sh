#!/bin/bash
# last validation of the committed binary: GPU tests, smoke, FP64 probe, 1e6 statistics table, bench
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2_final2_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/r2_final2_pytest_gpu.log
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2_final2_smoke.log 2>&1
tail -n 4 gpurun_out/r2_final2_smoke.log | cut -c1-200
python profiles/fp64_peak_probe.py > gpurun_out/r2_fp64_peak.json 2>/dev/null; cat gpurun_out/r2_fp64_peak.json
(time python profiles/suite_vs_reference.py suite_stats_ref_1e6.npz) > gpurun_out/r2_suite_vs_ref_1e6.txt 2>&1
tail -n 3 gpurun_out/r2_suite_vs_ref_1e6.txt | cut -c1-600
(time python bench.py) > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_final2_bench.json') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['suite']['wall_s'], d['roofline']['frac'], [s['evals_per_s'] for s in d['shapes']])
PY
