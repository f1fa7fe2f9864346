# Same-box A/B of the previous build (build/ab/libsdaopt_b200_prev.so) against the current one:
# headline shape and the Sutton-Chen shapes; then the GPU tests on the current build.
mkdir -p gpurun_out; rm -f gpurun_out/ab_s3.log
for lib in build/ab/libsdaopt_b200_prev.so sdaopt_b200/libsdaopt_b200.so; do
  echo "== $lib" >> gpurun_out/ab_s3.log
  SDA_LIB_PATH=$PWD/$lib python profiles/prof_run.py --chains 65536 --maxiter 300 --reps 2 >> gpurun_out/ab_s3.log 2>&1
done
echo "== current build, config 4" >> gpurun_out/ab_s3.log
CONFIG4_ONLY=1 timeout 200 python profiles/config34_run.py >> gpurun_out/ab_s3.log 2>&1
cat gpurun_out/ab_s3.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s3.log
tail -n 3 gpurun_out/pytest_gpu_s3.log
