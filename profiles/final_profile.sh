# Final evidence for the round (1 x B200): bench both arms, then the ncu launch list of the bench
# command, then one full capture of the dominant kernel.  Numbers printed under ncu are not bench values.
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/bench_ref_s2.json 2> gpurun_out/bench_ref_s2.err; cut -c1-260 gpurun_out/bench_ref_s2.json
python bench.py > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; cut -c1-330 gpurun_out/bench_s2.json
python bench.py --steps 1 --warmup 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_s2.csv \
    python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launches_s2.log 2>&1
SDA_PHASED=1 python profiles/prof_run.py --chains 65536 --maxiter 200 --reps 1 && \
SDA_PHASED=1 ncu --set full --clock-control none --import-source on -k regex:sda_anneal -s 20 -c 1 \
    -o gpurun_out/prof_phA_s2 -f python profiles/prof_run.py --chains 65536 --maxiter 200 --reps 1 > gpurun_out/ncu_phA_s2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -n 2
