#!/bin/bash
# round 2, GPU call 8: slab marked as global memory for the compiler; ncu of the 10x pile (the shallow cliff) and of the 40x pile, symbols saved with the captures
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/r2h_pytest_gpu.log
cuobjdump -elf daccord_b200/_build/libdaccord_b200.so | grep '\$_ZN' > gpurun_out/r2h_symbols.txt
ab() { local name=$1 mb=$2 cov=$3; shift 3; env "$@" timeout 600 python bench.py --mb $mb --coverage $cov --steps 3 --warmup 3 --cpu-sample-s 0 --cli 0 --truth-reads 0 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('$name value %.3f e2e %.3f parts %s' % (l['value']/1e6, l['e2e']['value']/1e6, l['e2e'].get('rank0_ms_per_step')))"; }
ab base40 10 40 X=1
ab w32 10 40 DCU_WPB=32 DCU_SYNC_GROUP=32
ab base10 5 10 X=1
ab w32_10 5 10 DCU_WPB=32 DCU_SYNC_GROUP=32
for cov in 10 40; do
  mb=2; [ $cov = 10 ] && mb=1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcu_window -s 2 -c 1 -f -o gpurun_out/r2h_cov${cov}_full python tools/ncu_target.py $mb $cov 2 > gpurun_out/r2h_ncu_cov$cov.log 2>&1; echo "ncu cov$cov rc=$?"; grep coverage gpurun_out/r2h_ncu_cov$cov.log | head -1
  ncu -i gpurun_out/r2h_cov${cov}_full.ncu-rep --page raw --csv > gpurun_out/r2h_cov${cov}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2h_cov${cov}_full.ncu-rep --page source --csv > gpurun_out/r2h_cov${cov}_source.csv 2>/dev/null
  rm -f gpurun_out/r2h_cov${cov}_full.ncu-rep
done
