#!/bin/bash
mkdir -p gpurun_out
python profiles/record_parity_pins.py > gpurun_out/r2_pins2.log 2>&1
export SDA_CONT_NS=148
python profiles/suite_contention.py 20000 Bukin06 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sda_anneal -s 1 -c 1 -f -o gpurun_out/r2_suite_bukin06 python profiles/suite_contention.py 20000 Bukin06 > gpurun_out/ncu1.log 2>&1
du -sh gpurun_out
