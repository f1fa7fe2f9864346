#!/bin/bash
export SDA_CONT_NS=13
timeout 600 python profiles/suite_contention.py 100000 Thurber DeVilliersGlasser02 Cola Meyer PowerSum Csendes Mishra06 NewFunction02 NewFunction01 Damavandi Weierstrass Deceptive CrownedCross CrossLegTable Mishra04 Shubert03 Zimmerman Bukin06 Stochastic XinSheYang01 Mishra03 2>&1 | cut -c1-125 > gpurun_out/r2_long_jobs_alone.log
cat gpurun_out/r2_long_jobs_alone.log
