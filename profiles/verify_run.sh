# GPU tests + suite-vs-reference table (1 x B200)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2.log
tail -n 3 gpurun_out/pytest_gpu_s2.log
python profiles/suite_vs_reference.py > gpurun_out/suite_vs_ref_s2.txt 2>&1; tail -n 1 gpurun_out/suite_vs_ref_s2.txt
grep -E "_20 |_30 |_50 |_100 |differs" gpurun_out/suite_vs_ref_s2.txt
