set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2.log
tail -n 3 gpurun_out/pytest_gpu_s2.log
for cfg in "173 8" "200 8" "230 8" "140 12" "170 12" "120 16"; do
  set -- $cfg
  echo "== SDA_ROUNDS=$1 SDA_STEPS_PER_ROUND=$2" >> gpurun_out/rounds_sweep.log
  SDA_ROUNDS=$1 SDA_STEPS_PER_ROUND=$2 python profiles/prof_run.py --chains 65536 --maxiter 1000 --reps 2 >> gpurun_out/rounds_sweep.log 2>&1
done
cat gpurun_out/rounds_sweep.log
