#!/bin/bash
mkdir -p gpurun_out
export SDA_CONT_NS=2368
python profiles/suite_contention.py 20000 Bukin06 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sda_anneal -s 1 -c 1 -f -o gpurun_out/r2_suite_bukin06_full python profiles/suite_contention.py 20000 Bukin06 > gpurun_out/ncu1.log 2>&1
du -sh gpurun_out
