#!/bin/bash
# round 2, GPU call 9: hybrid first pass (k-mer table in shared memory at 32 warps per SM) against the plain HBM first pass; barrier masks on the bench harness
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 gpurun_out/r2i_pytest_gpu.log
cuobjdump -elf daccord_b200/_build/libdaccord_b200.so | grep '\$_ZN' > gpurun_out/r2i_symbols.txt
ab() { local name=$1 mb=$2 cov=$3; shift 3; env "$@" timeout 600 python bench.py --mb $mb --coverage $cov --steps 3 --warmup 3 --cpu-sample-s 0 --cli 0 --truth-reads 0 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('$name value %.3f e2e %.3f second %d hard %d smem %s' % (l['value']/1e6, l['e2e']['value']/1e6, l['second_pass_windows'], l['hard_windows'], l['smem_pass']))"; }
ab hybrid40 10 40 X=1
ab plain40 10 40 DCU_HYBRID=0
ab hybrid40_m0 10 40 DCU_SYNC_MASK=0
ab plain40_m0 10 40 DCU_HYBRID=0 DCU_SYNC_MASK=0
ab plain40_m8 10 40 DCU_HYBRID=0 DCU_SYNC_MASK=8
ab hybrid30 10 30 X=1
ab plain30 10 30 DCU_HYBRID=0
ab hybrid20 10 20 DCU_HYBRID_MIN_DEPTH=10
ab plain20 10 20 X=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcuh_window -s 1 -c 1 -f -o gpurun_out/r2i_hyb_full python tools/ncu_target.py 2 40 2 > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"; grep coverage gpurun_out/r2i_ncu.log | head -1
ncu -i gpurun_out/r2i_hyb_full.ncu-rep --page raw --csv > gpurun_out/r2i_hyb_raw.csv 2>/dev/null
ncu -i gpurun_out/r2i_hyb_full.ncu-rep --page source --csv > gpurun_out/r2i_hyb_source.csv 2>/dev/null
rm -f gpurun_out/r2i_hyb_full.ncu-rep
