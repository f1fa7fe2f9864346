# Round-end measurement set (1 x B200): A/B against the previous build, GPU tests, suite-vs-reference
# table, the 250-job suite, bench both arms.
mkdir -p gpurun_out
rm -f gpurun_out/ab_s2.log
if [ -f build/ab/libsdaopt_b200_prev.so ]; then
  for lib in build/ab/libsdaopt_b200_prev.so sdaopt_b200/libsdaopt_b200.so build/ab/libsdaopt_b200_prev.so sdaopt_b200/libsdaopt_b200.so; do
    echo "== $lib" >> gpurun_out/ab_s2.log
    SDA_LIB_PATH=$PWD/$lib python profiles/prof_run.py --chains 65536 --maxiter 300 --reps 2 >> gpurun_out/ab_s2.log 2>&1
  done
  cat gpurun_out/ab_s2.log
fi
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2.log
tail -n 3 gpurun_out/pytest_gpu_s2.log
python profiles/suite_vs_reference.py > gpurun_out/suite_vs_ref_s2.txt 2>&1; tail -n 2 gpurun_out/suite_vs_ref_s2.txt
grep -E "Stochastic|XinSheYang01 " gpurun_out/suite_vs_ref_s2.txt
rm -rf gpurun_out/DATA_s2
( time python -m sdaopt_b200.benchmark.workflow --nb-runs 100 --output-folder gpurun_out/DATA_s2 ) > gpurun_out/suite_s2.log 2>&1; tail -n 5 gpurun_out/suite_s2.log
rm -f gpurun_out/DATA_s2/*.data
if [ "$1" = "bench" ]; then
python bench.py --impl reference > gpurun_out/bench_ref_s2.json 2> gpurun_out/bench_ref_s2.err; cat gpurun_out/bench_ref_s2.json | cut -c1-300
python bench.py > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; cat gpurun_out/bench_s2.json | cut -c1-400
fi
