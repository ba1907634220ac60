#!/bin/bash
# round 2, GPU call 15: ncu capture of the 10x pile (the shallow cliff) for the record
set -u
mkdir -p gpurun_out
cuobjdump -elf daccord_b200/_build/libdaccord_b200.so | grep '\$_ZN' > gpurun_out/r2o_symbols.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcu_window -s 1 -c 1 -f -o gpurun_out/r2o_cov10_full python tools/ncu_target.py 1 10 2 > gpurun_out/r2o_ncu_cov10.log 2>&1; echo "ncu rc=$?"; grep coverage gpurun_out/r2o_ncu_cov10.log | head -2
ncu -i gpurun_out/r2o_cov10_full.ncu-rep --page raw --csv > gpurun_out/r2o_cov10_raw.csv 2>/dev/null
ncu -i gpurun_out/r2o_cov10_full.ncu-rep --page source --csv > gpurun_out/r2o_cov10_source.csv 2>/dev/null
rm -f gpurun_out/r2o_cov10_full.ncu-rep
